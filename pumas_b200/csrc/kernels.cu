/*
 * kernels.cu -- sm_100a transport kernels.
 *
 * transport_wavefront_kernel (the product path)
 *   A persistent kernel: a few CTAs per SM, each owning a pool of SLOTS track
 *   slots parked in shared memory (struct-of-arrays by 64 bit word, so that
 *   thread t touching slot t is bank conflict free). Every round the CTA
 *   regroups its slots by pending work (the phases of pb_transport.h):
 *     - slots that need a Monte Carlo step are stepped, one thread each, by
 *       full warps that all run the same code (step_transport);
 *     - slots with a pending hard elastic event (1 step in 6) are compacted
 *       and handled together; slots with a pending hard inelastic event
 *       (1 step in 340, but thousands of instructions: DCS models, Brent
 *       maximiser) are left parked until enough of them have accumulated;
 *     - finished slots are flushed to HBM and refilled from a global atomic
 *       cursor, which also removes the tail of long tracks.
 *   This is the wavefront / event-queue scheme of BASELINE.json's north_star
 *   at CTA scope: regrouping happens through shared memory, no global queues.
 *   A monolithic thread-per-track loop spends 2/3 of its issue slots on
 *   divergence (profiles/r1_v0_monolithic.md: 10.4 of 32 lanes active).
 *
 * transport_kernel (reference kernel, kept for A/B parity and profiling)
 *   One thread owns one track and runs its whole step loop.
 *
 * The physics tables stay in HBM / L2 (1.6 MB, L2 resident) and are read
 * through the read-only path; a track state crosses HBM twice (104 B in,
 * 104 B out).
 */
#include <cuda_runtime.h>

#include "kernels.h"
#include "pb_transport.h"

namespace pb {

/* Philox4x32-10 stream of one track, register resident. Same stream as
 * philox.h (which the CPU oracle uses); the four words of a block are kept in
 * scalar registers and selected without indexing. */
struct DeviceRng {
        uint32_t key0, key1, trk0, trk1;
        uint32_t block, have, count;
        uint32_t o0, o1, o2, o3;
        uint32_t pad_;

        __device__ __forceinline__ void init(uint64_t seed, uint64_t track)
        {
                key0 = (uint32_t)seed;
                key1 = (uint32_t)(seed >> 32);
                trk0 = (uint32_t)track;
                trk1 = (uint32_t)(track >> 32);
                block = have = count = 0;
                o0 = o1 = o2 = o3 = 0;
                pad_ = 0;
        }

        __device__ __forceinline__ void refill()
        {
                uint32_t c0 = block, c1 = 0, c2 = trk0, c3 = trk1;
                uint32_t k0 = key0, k1 = key1;
#pragma unroll
                for (int r = 0; r < 10; r++) {
                        const uint32_t hi0 = __umulhi(0xD2511F53u, c0);
                        const uint32_t lo0 = 0xD2511F53u * c0;
                        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2);
                        const uint32_t lo1 = 0xCD9E8D57u * c2;
                        c0 = hi1 ^ c1 ^ k0;
                        c1 = lo1;
                        c2 = hi0 ^ c3 ^ k1;
                        c3 = lo0;
                        k0 += 0x9E3779B9u;
                        k1 += 0xBB67AE85u;
                }
                o0 = c0;
                o1 = c1;
                o2 = c2;
                o3 = c3;
                block++;
                have = 4;
        }

        __device__ __forceinline__ double uniform()
        {
                if (have == 0) refill();
                const uint32_t w = (have == 4) ? o0 :
                    (have == 3)                ? o1 :
                    (have == 2)                ? o2 :
                                                 o3;
                have--;
                count++;
                return ((double)w + 0.5) * (1.0 / 4294967296.0);
        }
};

/* ------------------------------------------------------------------------
 * Reference kernel: one thread, one track, whole loop.
 * ------------------------------------------------------------------------ */
__global__ void __launch_bounds__(TRANSPORT_BLOCK)
    transport_kernel(const Physics p, const Settings cfg, const Geometry geo,
        const uint64_t seed, const uint64_t offset, const long n,
        State * __restrict__ states, int * __restrict__ events,
        int * __restrict__ media, int * __restrict__ status,
        unsigned int * __restrict__ n_random, unsigned int * __restrict__ n_steps)
{
        const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= n) return;

        State state = states[i];
        DeviceRng rng;
        rng.init(seed, offset + (uint64_t)i);
        unsigned int steps = 0;
        const Outcome out = transport_track(p, cfg, geo, rng, state, &steps);
        states[i] = state;
        if (events != nullptr) events[i] = out.event;
        if (media != nullptr) {
                media[2 * i] = out.medium[0];
                media[2 * i + 1] = out.medium[1];
        }
        if (out.rc != RETURN_SUCCESS) atomicCAS(status, 0, out.rc);
        if (n_random != nullptr) n_random[i] = rng.count;
        if (n_steps != nullptr) n_steps[i] = steps;
}

/* ------------------------------------------------------------------------
 * Wavefront kernel.
 * ------------------------------------------------------------------------ */
constexpr int WF_THREADS = 128;
constexpr int WF_SLOTS = 160;
constexpr int WF_CTAS_PER_SM = 3;
/* A parked hard inelastic event is processed once this many are pending (or
 * when the CTA has nothing else to run). */
constexpr int WF_DEL_THRESHOLD = 16;
constexpr int PHASE_EMPTY = 4;
constexpr int N_PHASES = 5;

/* Everything a parked track needs, as 64 bit words. */
struct Slot {
        State state;
        Track trk;
        Loop lp;
        DeviceRng rng;
        long index;   /* position in the batch */
        int medium0;  /* first medium (media[0] of the reference) */
        int pad_;
};
static_assert(sizeof(Slot) % 8 == 0, "Slot must be a whole number of words");
constexpr int SLOT_WORDS = sizeof(Slot) / 8;

struct WavefrontShared {
        unsigned long long words[SLOT_WORDS * WF_SLOTS];
        int phase[WF_SLOTS];
        unsigned short list[N_PHASES][WF_SLOTS];
        int count[N_PHASES];
        long refill_base;
        int exhausted;
};

__device__ __forceinline__ void slot_load(
    Slot & s, const unsigned long long * words, int slot)
{
        /* (memcpy, not a pointer cast: the members are read back through
         * their own types) */
        unsigned long long w[SLOT_WORDS];
#pragma unroll
        for (int k = 0; k < SLOT_WORDS; k++) w[k] = words[k * WF_SLOTS + slot];
        memcpy(&s, w, sizeof(Slot));
}

__device__ __forceinline__ void slot_store(
    const Slot & s, unsigned long long * words, int slot)
{
        unsigned long long w[SLOT_WORDS];
        memcpy(w, &s, sizeof(Slot));
#pragma unroll
        for (int k = 0; k < SLOT_WORDS; k++) words[k * WF_SLOTS + slot] = w[k];
}

/* Write a finished track back to HBM. */
__device__ __forceinline__ void track_flush(const Slot & s, const Outcome & out,
    State * __restrict__ states, int * __restrict__ events,
    int * __restrict__ media, int * __restrict__ status,
    unsigned int * __restrict__ n_random, unsigned int * __restrict__ n_steps)
{
        const long i = s.index;
        states[i] = s.state;
        if (events != nullptr) events[i] = out.event;
        if (media != nullptr) {
                media[2 * i] = out.medium[0];
                media[2 * i + 1] = out.medium[1];
        }
        if (out.rc != RETURN_SUCCESS) atomicCAS(status, 0, out.rc);
        if (n_random != nullptr) n_random[i] = s.rng.count;
        if (n_steps != nullptr) n_steps[i] = s.trk.n_steps;
}

__global__ void __launch_bounds__(WF_THREADS, WF_CTAS_PER_SM)
    transport_wavefront_kernel(const Physics p, const Settings cfg,
        const Geometry geo, const uint64_t seed, const uint64_t offset,
        const long n, State * __restrict__ states, int * __restrict__ events,
        int * __restrict__ media, int * __restrict__ status,
        unsigned int * __restrict__ n_random,
        unsigned int * __restrict__ n_steps,
        unsigned long long * __restrict__ cursor)
{
        extern __shared__ __align__(16) unsigned char smem_raw[];
        WavefrontShared & sh = *reinterpret_cast<WavefrontShared *>(smem_raw);
        const int tid = threadIdx.x;

        for (int s = tid; s < WF_SLOTS; s += WF_THREADS) sh.phase[s] = PHASE_EMPTY;
        if (tid == 0) sh.exhausted = 0;
        __syncthreads();

        for (;;) {
                /* (read before the barriers below: thread 0 updates it after
                 * them, in the refill) */
                const int exhausted = sh.exhausted;

                /* ---- regroup the slots by pending work ------------------ */
                if (tid < N_PHASES) sh.count[tid] = 0;
                __syncthreads();
                for (int s = tid; s < WF_SLOTS; s += WF_THREADS) {
                        const int ph = sh.phase[s];
                        const int pos = atomicAdd(&sh.count[ph], 1);
                        sh.list[ph][pos] = (unsigned short)s;
                }
                __syncthreads();
                const int n_step = sh.count[PHASE_STEP];
                const int n_ehs = sh.count[PHASE_EHS];
                const int n_del = sh.count[PHASE_DEL];
                const int n_empty = sh.count[PHASE_EMPTY];
                if ((n_empty == WF_SLOTS) && exhausted) break;

                /* ---- refill the empty slots from the global cursor ------ */
                if ((n_empty > 0) && !exhausted) {
                        if (tid == 0) {
                                const long base = (long)atomicAdd(
                                    cursor, (unsigned long long)n_empty);
                                sh.refill_base = base;
                                if (base + n_empty >= n) sh.exhausted = 1;
                        }
                        __syncthreads();
                        const long base = sh.refill_base;
                        for (int j = tid; j < n_empty; j += WF_THREADS) {
                                const long i = base + j;
                                if (i >= n) continue;
                                const int slot = sh.list[PHASE_EMPTY][j];
                                Slot s;
                                s.index = i;
                                s.state = states[i];
                                s.rng.init(seed, offset + (uint64_t)i);
                                s.pad_ = 0;
                                Outcome out;
                                const int ph = track_begin(p, cfg, geo, s.rng,
                                    s.state, s.trk, s.lp, out);
                                s.medium0 = out.medium[0];
                                if (ph == PHASE_DONE) {
                                        track_flush(s, out, states, events,
                                            media, status, n_random, n_steps);
                                } else {
                                        slot_store(s, sh.words, slot);
                                        sh.phase[slot] = ph;
                                }
                        }
                        /* The new slots join the lists at the next round. */
                }

                /* ---- Monte Carlo steps ------------------------------------ */
                if (tid < n_step) {
                        const int slot = sh.list[PHASE_STEP][tid];
                        Slot s;
                        slot_load(s, sh.words, slot);
                        int ph = stepping_step(
                            p, cfg, geo, s.trk, s.rng, s.state, s.lp);
                        if (ph == PHASE_DONE) {
                                Outcome out;
                                out.medium[0] = s.medium0;
                                track_end(s.trk, s.lp, out);
                                track_flush(s, out, states, events, media,
                                    status, n_random, n_steps);
                                ph = PHASE_EMPTY;
                        } else {
                                slot_store(s, sh.words, slot);
                        }
                        sh.phase[slot] = ph;
                }

                /* ---- hard elastic events: handled every round ------------- */
                for (int j = tid; j < n_ehs; j += WF_THREADS) {
                        const int slot = sh.list[PHASE_EHS][j];
                        Slot s;
                        slot_load(s, sh.words, slot);
                        int ph = stepping_ehs(
                            p, cfg, geo, s.trk, s.rng, s.state, s.lp);
                        if (ph == PHASE_DONE) {
                                Outcome out;
                                out.medium[0] = s.medium0;
                                track_end(s.trk, s.lp, out);
                                track_flush(s, out, states, events, media,
                                    status, n_random, n_steps);
                                ph = PHASE_EMPTY;
                        } else {
                                slot_store(s, sh.words, slot);
                        }
                        sh.phase[slot] = ph;
                }

                /* ---- hard inelastic events: parked until worth a pass ----- */
                if ((n_del >= WF_DEL_THRESHOLD) ||
                    ((n_del > 0) && (n_step == 0) && (n_ehs == 0))) {
                        /* served from the top so that they land on other
                         * warps than the elastic events */
                        for (int j = WF_THREADS - 1 - tid; j < n_del;
                             j += WF_THREADS) {
                                const int slot = sh.list[PHASE_DEL][j];
                                Slot s;
                                slot_load(s, sh.words, slot);
                                int ph = stepping_del(
                                    p, cfg, geo, s.trk, s.rng, s.state, s.lp);
                                if (ph == PHASE_DONE) {
                                        Outcome out;
                                        out.medium[0] = s.medium0;
                                        track_end(s.trk, s.lp, out);
                                        track_flush(s, out, states, events,
                                            media, status, n_random, n_steps);
                                        ph = PHASE_EMPTY;
                                } else {
                                        slot_store(s, sh.words, slot);
                                }
                                sh.phase[slot] = ph;
                        }
                }
                __syncthreads();
        }
}

static int g_wavefront_grid = 0;

cudaError_t launch_transport(const Physics & p, const Settings & cfg,
    const Geometry & geo, uint64_t seed, uint64_t offset, long n,
    State * states, int * events, int * media, int * status,
    unsigned int * n_random, unsigned int * n_steps,
    unsigned long long * cursor, int kernel, cudaStream_t stream)
{
        if (n <= 0) return cudaSuccess;
        if (kernel == KERNEL_MONOLITHIC) {
                const long blocks = (n + TRANSPORT_BLOCK - 1) / TRANSPORT_BLOCK;
                transport_kernel<<<(unsigned int)blocks, TRANSPORT_BLOCK, 0,
                    stream>>>(p, cfg, geo, seed, offset, n, states, events,
                    media, status, n_random, n_steps);
                return cudaGetLastError();
        }

        const size_t smem = sizeof(WavefrontShared);
        if (g_wavefront_grid == 0) {
                cudaError_t e = cudaFuncSetAttribute(transport_wavefront_kernel,
                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
                int dev = 0, sms = 0, per_sm = 0;
                cudaGetDevice(&dev);
                cudaDeviceGetAttribute(
                    &sms, cudaDevAttrMultiProcessorCount, dev);
                e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm,
                    transport_wavefront_kernel, WF_THREADS, smem);
                if (e != cudaSuccess) return e;
                if (per_sm < 1) per_sm = 1;
                g_wavefront_grid = sms * per_sm;
        }
        /* Persistent grid: one wave of resident CTAs, fewer for small n. */
        long want = (n + WF_SLOTS - 1) / WF_SLOTS;
        int grid = (want < g_wavefront_grid) ? (int)want : g_wavefront_grid;
        cudaError_t e =
            cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return e;
        transport_wavefront_kernel<<<grid, WF_THREADS, smem, stream>>>(p, cfg,
            geo, seed, offset, n, states, events, media, status, n_random,
            n_steps, cursor);
        return cudaGetLastError();
}

} /* namespace pb */
