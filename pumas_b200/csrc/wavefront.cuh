/*
 * wavefront.cuh -- the wavefront transport kernel (template) and its launcher.
 *
 * A persistent kernel: every warp owns a pool of `slots` parked tracks (packed
 * in 35 words of 64 bits each, in a global memory workspace that stays L2
 * resident; loads and stores are coalesced) and, every round, regroups them by
 * pending work, the phases of pb_transport.h:
 *   - up to 32 tracks that need a Monte Carlo step are stepped, one lane each,
 *     all lanes running step_transport together;
 *   - tracks with a pending hard elastic event (1 step in 6) or hard inelastic
 *     event (1 step in 340, but thousands of instructions) stay parked until
 *     enough of them have accumulated to fill a pass;
 *   - finished tracks are flushed to HBM and their slots refilled from a
 *     global atomic cursor, which also removes the tail of long tracks.
 * This is the wavefront / event-queue scheme of BASELINE.json's north_star at
 * warp scope: ballots build the work lists, no block barrier is involved, a
 * slow pass (the photonuclear Brent maximiser) only delays its own warp.
 *
 * The kernel is a template on the settings and geometry types so that the
 * common configurations get kernels stripped of the code they cannot reach
 * (pb_common.h: StaticSettings, pb_media.h: UniformGeometry): the general
 * kernel is ~45 k instructions, several times the instruction cache, and
 * `no_instruction` was its first stall reason (profiles/r1_v2).
 */
#ifndef PB_WAVEFRONT_CUH
#define PB_WAVEFRONT_CUH

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "pb_transport.h"

namespace pb {

/* One Philox4x32-10 block, out of line (one copy instead of one per call site
 * of uniform()). */
static __device__ __noinline__ uint4 philox_block(
    uint32_t key0, uint32_t key1, uint32_t block, uint32_t trk0, uint32_t trk1)
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
        return make_uint4(c0, c1, c2, c3);
}

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
                const uint4 o = philox_block(key0, key1, block, trk0, trk1);
                o0 = o.x;
                o1 = o.y;
                o2 = o.z;
                o3 = o.w;
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
 * Wavefront kernel.
 * ------------------------------------------------------------------------ */
/* One CTA per SM. Its warps are `WF_WARPS - WF_DEL_WARPS` stepping warps, kept
 * in lockstep by a named barrier so that they walk the (large) stepping code
 * together and share its instruction cache lines, plus WF_DEL_WARPS warps that
 * serve the hard inelastic events of the whole CTA asynchronously. */
#ifndef PB_WF_WARPS
#define PB_WF_WARPS 16
#endif
#ifndef PB_WF_DEL_WARPS
#define PB_WF_DEL_WARPS 1
#endif
constexpr int WF_WARPS = PB_WF_WARPS;
constexpr int WF_DEL_WARPS = PB_WF_DEL_WARPS;
constexpr int WF_STEP_WARPS = WF_WARPS - WF_DEL_WARPS;
constexpr int WF_THREADS = 32 * WF_WARPS;
constexpr int WF_SLOTS_MAX = 128; /* track slots per warp (upper bound) */
#ifndef PB_WF_CTAS_PER_SM
#define PB_WF_CTAS_PER_SM 1
#endif
constexpr int WF_CTAS_PER_SM = PB_WF_CTAS_PER_SM;
/* Parked events are served once this many are pending, or earlier when the
 * warp cannot fill a stepping pass. */
/* The stepping warps re-align (named barrier) every WF_SYNC_ROUNDS rounds: often
 * enough to keep sharing instruction cache lines, seldom enough to average out
 * the round to round variance of the slowest warp. */
#ifndef PB_WF_SYNC_ROUNDS
#define PB_WF_SYNC_ROUNDS 1
#endif
constexpr int WF_SYNC_ROUNDS = PB_WF_SYNC_ROUNDS;
constexpr int WF_SLOTS_DEFAULT = 96;
constexpr int WF_EHS_THRESHOLD = 16;
constexpr int WF_EHS_PERIOD = 4;
constexpr int WF_DEL_THRESHOLD = 8;
/* Pool phases: PHASE_STEP, PHASE_EHS, PHASE_DEL (pb_transport.h), then */
constexpr int PHASE_DELQ = 3;  /* hard inelastic event handed to a DEL warp */
constexpr int PHASE_EMPTY = 4;
constexpr int N_PHASES = 5;

/* A parked track, packed in 35 words of 64 bits. */
constexpr int SLOT_WORDS = 35;
struct Slot {
        State state;
        Track trk;
        Loop lp;
        DeviceRng rng;
        long index;  /* position in the batch */
        int medium0; /* first medium (media[0] of the reference) */
};

/* Per warp pool in shared memory: `slots` packed tracks, their phases and the
 * per phase work lists. */
struct WarpPool {
        unsigned long long * words; /* [SLOT_WORDS][slots] */
        unsigned char * phase;      /* [slots] */
        unsigned char * list;       /* [N_PHASES][slots] */
        int slots;
};

/* Shared memory per warp: phases and work lists. The packed tracks themselves
 * live in a global memory workspace (L2 resident, coalesced: 35 words per
 * track and per pass), which leaves the whole 228 KB of the SM to L1 for the
 * physics tables. */
__host__ __device__ inline size_t warp_pool_bytes(int slots)
{
        return (((size_t)slots * (1 + N_PHASES)) + 15) & ~(size_t)15;
}

__host__ __device__ inline size_t warp_slot_words(int slots)
{
        return (size_t)slots * SLOT_WORDS;
}

__device__ __forceinline__ unsigned long long pack_d(double x)
{
        return (unsigned long long)__double_as_longlong(x);
}

__device__ __forceinline__ double unpack_d(unsigned long long w)
{
        return __longlong_as_double((long long)w);
}

__device__ __forceinline__ unsigned long long pack_2(uint32_t lo, uint32_t hi)
{
        return (unsigned long long)lo | ((unsigned long long)hi << 32);
}

/* Word w of slot s lives at words[w * WF_SLOTS + s]: lanes that touch
 * consecutive slots hit consecutive banks. */
/* (accessed with .cg: the DEL warps and the owner warp exchange slots through
 * L2, whatever L1 holds) */
#define SLOT_ST(k, v) __stcg(&words[(k) * n_slots + slot], (v))
#define SLOT_W(k) __ldcg(&words[(k) * n_slots + slot])

__device__ __forceinline__ void slot_store(const Slot & s,
    unsigned long long * words, int n_slots, int slot, uint64_t seed)
{
        (void)seed;
        const State & st = s.state;
        SLOT_ST(0, pack_d(st.charge));
        SLOT_ST(1, pack_d(st.energy));
        SLOT_ST(2, pack_d(st.distance));
        SLOT_ST(3, pack_d(st.grammage));
        SLOT_ST(4, pack_d(st.time));
        SLOT_ST(5, pack_d(st.weight));
        SLOT_ST(6, pack_d(st.position[0]));
        SLOT_ST(7, pack_d(st.position[1]));
        SLOT_ST(8, pack_d(st.position[2]));
        SLOT_ST(9, pack_d(st.direction[0]));
        SLOT_ST(10, pack_d(st.direction[1]));
        SLOT_ST(11, pack_d(st.direction[2]));
        const Track & t = s.trk;
        SLOT_ST(12, pack_d(t.lifetime));
        SLOT_ST(13, pack_d(t.step_X_limit));
        SLOT_ST(14, pack_d(t.step_invlb1));
        SLOT_ST(15, pack_d(t.randn_next));
        const Loop & l = s.lp;
        SLOT_ST(16, pack_d(l.ti));
        SLOT_ST(17, pack_d(l.wi));
        SLOT_ST(18, pack_d(l.Xi));
        SLOT_ST(19, pack_d(l.Xf));
        SLOT_ST(20, pack_d(l.dei));
        SLOT_ST(21, pack_d(l.grammage_max));
        SLOT_ST(22, pack_d(l.step_max_medium));
        SLOT_ST(23, pack_d(l.step_max_locals));
        SLOT_ST(24, pack_d(l.locals.density));
        SLOT_ST(25, pack_d(l.locals.magnet[0]));
        SLOT_ST(26, pack_d(l.locals.magnet[1]));
        SLOT_ST(27, pack_d(l.locals.magnet[2]));
        /* flags: events are 14 bit masks, media and materials are small */
        const uint32_t f0 = (uint32_t)t.step_event |
            ((uint32_t)t.step_foreseen << 14) |
            ((uint32_t)(t.step_first & 1) << 28) |
            ((uint32_t)(t.randn_done & 1) << 29) |
            ((uint32_t)(st.decayed & 1) << 30) |
            ((uint32_t)(l.locals.magnetized & 1) << 31);
        const uint32_t f1 = (uint32_t)(l.medium & 0xff) |
            ((uint32_t)(l.material & 0xff) << 8) |
            ((uint32_t)(s.medium0 & 0xff) << 16) |
            ((uint32_t)(l.straight & 1) << 24) |
            ((uint32_t)(l.step_max_type & 1) << 25) |
            ((uint32_t)(l.first & 1) << 26) |
            ((uint32_t)(t.error & 0x1f) << 27);
        SLOT_ST(28, pack_2(f0, f1));
        SLOT_ST(29, pack_2(t.n_steps, s.rng.count));
        SLOT_ST(30, pack_2(s.rng.block, s.rng.have));
        SLOT_ST(31, pack_2(s.rng.o0, s.rng.o1));
        SLOT_ST(32, pack_2(s.rng.o2, s.rng.o3));
        SLOT_ST(33, (unsigned long long)s.index |
            ((unsigned long long)(t.iK & 0xff) << 56));
        SLOT_ST(34, pack_2(s.rng.trk0, s.rng.trk1));
}

__device__ __forceinline__ void slot_load(Slot & s,
    const unsigned long long * words, int n_slots, int slot, uint64_t seed)
{
        State & st = s.state;
        st.charge = unpack_d(SLOT_W(0));
        st.energy = unpack_d(SLOT_W(1));
        st.distance = unpack_d(SLOT_W(2));
        st.grammage = unpack_d(SLOT_W(3));
        st.time = unpack_d(SLOT_W(4));
        st.weight = unpack_d(SLOT_W(5));
        st.position[0] = unpack_d(SLOT_W(6));
        st.position[1] = unpack_d(SLOT_W(7));
        st.position[2] = unpack_d(SLOT_W(8));
        st.direction[0] = unpack_d(SLOT_W(9));
        st.direction[1] = unpack_d(SLOT_W(10));
        st.direction[2] = unpack_d(SLOT_W(11));
        Track & t = s.trk;
        t.lifetime = unpack_d(SLOT_W(12));
        t.step_X_limit = unpack_d(SLOT_W(13));
        t.step_invlb1 = unpack_d(SLOT_W(14));
        t.randn_next = unpack_d(SLOT_W(15));
        Loop & l = s.lp;
        l.ti = unpack_d(SLOT_W(16));
        l.wi = unpack_d(SLOT_W(17));
        l.Xi = unpack_d(SLOT_W(18));
        l.Xf = unpack_d(SLOT_W(19));
        l.dei = unpack_d(SLOT_W(20));
        l.grammage_max = unpack_d(SLOT_W(21));
        l.step_max_medium = unpack_d(SLOT_W(22));
        l.step_max_locals = unpack_d(SLOT_W(23));
        l.locals.density = unpack_d(SLOT_W(24));
        l.locals.magnet[0] = unpack_d(SLOT_W(25));
        l.locals.magnet[1] = unpack_d(SLOT_W(26));
        l.locals.magnet[2] = unpack_d(SLOT_W(27));
        const unsigned long long f = SLOT_W(28);
        const uint32_t f0 = (uint32_t)f, f1 = (uint32_t)(f >> 32);
        t.step_event = (int)(f0 & 0x3fff);
        t.step_foreseen = (int)((f0 >> 14) & 0x3fff);
        t.step_first = (int)((f0 >> 28) & 1);
        t.randn_done = (int)((f0 >> 29) & 1);
        st.decayed = (int)((f0 >> 30) & 1);
        l.locals.magnetized = (int)((f0 >> 31) & 1);
        l.medium = (int)(f1 & 0xff);
        l.material = (int)((f1 >> 8) & 0xff);
        s.medium0 = (int)((f1 >> 16) & 0xff);
        l.straight = (int)((f1 >> 24) & 1);
        l.step_max_type = (int)((f1 >> 25) & 1);
        l.first = (int)((f1 >> 26) & 1);
        t.error = (int)((f1 >> 27) & 0x1f);
        const unsigned long long c = SLOT_W(29);
        t.n_steps = (uint32_t)c;
        s.rng.count = (uint32_t)(c >> 32);
        const unsigned long long b = SLOT_W(30);
        s.rng.block = (uint32_t)b;
        s.rng.have = (uint32_t)(b >> 32);
        const unsigned long long o01 = SLOT_W(31), o23 = SLOT_W(32);
        s.rng.o0 = (uint32_t)o01;
        s.rng.o1 = (uint32_t)(o01 >> 32);
        s.rng.o2 = (uint32_t)o23;
        s.rng.o3 = (uint32_t)(o23 >> 32);
        const unsigned long long ix = SLOT_W(33);
        s.index = (long)(ix & 0x00ffffffffffffffull);
        t.iK = (int)(ix >> 56);
        const unsigned long long tr = SLOT_W(34);
        s.rng.trk0 = (uint32_t)tr;
        s.rng.trk1 = (uint32_t)(tr >> 32);
        s.rng.key0 = (uint32_t)seed;
        s.rng.key1 = (uint32_t)(seed >> 32);
        s.rng.pad_ = 0;
}
#undef SLOT_W
#undef SLOT_ST

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

struct WavefrontArgs {
        uint64_t seed, offset;
        long n;
        State * states;
        int * events;
        int * media;
        int * status;
        unsigned int * n_random;
        unsigned int * n_steps;
        unsigned long long * cursor;
        unsigned long long * slot_words; /* [warp][SLOT_WORDS][slots] */
        unsigned long long * stats; /* optional scheduler statistics */
        int slots;                  /* track slots per warp */
        int ehs_threshold;
        int ehs_period; /* > 0: serve the elastic events every that many rounds */
        int del_threshold;
};

/* Scheduler statistics (diagnostics), summed over warps. */
enum {
        STAT_ROUNDS = 0,
        STAT_STEP_PASSES,
        STAT_STEP_LANES,
        STAT_EHS_PASSES,
        STAT_EHS_LANES,
        STAT_DEL_PASSES,
        STAT_DEL_LANES,
        STAT_REFILLS,
        N_STATS
};

/* End of a pass for one slot: flush a finished track or park it again. */
__device__ __forceinline__ void slot_retire(Slot & s, int ph, WarpPool & pool,
    int slot, const WavefrontArgs & a)
{
        if (ph == PHASE_DONE) {
                Outcome out;
                out.medium[0] = (s.medium0 == 0xff) ? -1 : s.medium0;
                track_end(s.trk, s.lp, out);
                track_flush(s, out, a.states, a.events, a.media, a.status,
                    a.n_random, a.n_steps);
                ph = PHASE_EMPTY;
        } else {
                slot_store(s, pool.words, pool.slots, slot, a.seed);
        }
        /* the slot words must be visible before the phase is */
        __threadfence();
        ((volatile unsigned char *)pool.phase)[slot] = (unsigned char)ph;
}

/* CTA wide shared state. */
struct WavefrontShared {
        volatile unsigned int del_tail; /* events pushed so far */
        volatile unsigned int n_done;   /* stepping warps out of work */
        unsigned int pad_[2];
};

__device__ __forceinline__ void named_barrier_sync(int id, int threads)
{
        asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <class Cfg, class Geo>
__global__ void __launch_bounds__(WF_THREADS, WF_CTAS_PER_SM)
    transport_wavefront_kernel(const Physics p, const Cfg cfg, const Geo geo,
        const WavefrontArgs a)
{
        extern __shared__ __align__(16) unsigned char smem_raw[];
        const int lane = threadIdx.x & 31;
        const int warp = threadIdx.x >> 5;
        const int n_slots = a.slots;
        WavefrontShared & sh = *reinterpret_cast<WavefrontShared *>(smem_raw);
        /* queue of (warp << 8 | slot) entries, one per slot of the CTA */
        volatile unsigned short * const del_queue =
            reinterpret_cast<volatile unsigned short *>(
                smem_raw + sizeof(WavefrontShared));
        const int queue_size = WF_STEP_WARPS * n_slots;
        unsigned char * const pools = smem_raw + sizeof(WavefrontShared) +
            ((2 * (size_t)queue_size + 15) & ~(size_t)15);
        unsigned long long * const cta_words = a.slot_words +
            (size_t)blockIdx.x * WF_STEP_WARPS * warp_slot_words(n_slots);

        if (threadIdx.x == 0) {
                sh.del_tail = 0;
                sh.n_done = 0;
        }
        for (int s = threadIdx.x; s < WF_STEP_WARPS * n_slots; s += WF_THREADS) {
                (pools + (s / n_slots) * warp_pool_bytes(n_slots))[s % n_slots] =
                    (unsigned char)PHASE_EMPTY;
                del_queue[s] = 0xffffu;
        }
        __syncthreads();

        unsigned long long stat[N_STATS];
#pragma unroll
        for (int k = 0; k < N_STATS; k++) stat[k] = 0;

        if ((WF_DEL_WARPS > 0) && (warp >= WF_STEP_WARPS)) {
                /* ---- DEL warp: serve the queue of hard inelastic events -- */
                unsigned int head = 0; /* entries consumed so far */
                int idle = 0;
                for (;;) {
                        const unsigned int tail = sh.del_tail;
                        const int avail = (int)(tail - head);
                        const bool all_done = (sh.n_done == WF_STEP_WARPS);
                        if ((avail >= a.del_threshold) ||
                            ((avail > 0) && ((idle > 8) || all_done))) {
                                const int nb = (avail < 32) ? avail : 32;
                                if (lane < nb) {
                                        /* wait for the producer's entry */
                                        const unsigned int at =
                                            (head + lane) % queue_size;
                                        unsigned int e;
                                        do {
                                                e = del_queue[at];
                                        } while (e == 0xffffu);
                                        del_queue[at] = 0xffffu;
                                        __threadfence();
                                        const int w = (int)(e >> 8);
                                        const int slot = (int)(e & 0xff);
                                        WarpPool pool;
                                        pool.words = cta_words +
                                            w * warp_slot_words(n_slots);
                                        pool.phase = pools +
                                            w * warp_pool_bytes(n_slots);
                                        pool.list = nullptr;
                                        pool.slots = n_slots;
                                        Slot s;
                                        slot_load(s, pool.words, n_slots, slot,
                                            a.seed);
                                        const int ph = stepping_del(p, cfg, geo,
                                            s.trk, s.rng, s.state, s.lp);
                                        slot_retire(s, ph, pool, slot, a);
                                }
                                __syncwarp();
                                stat[STAT_DEL_PASSES]++;
                                stat[STAT_DEL_LANES] += nb;
                                head += nb;
                                idle = 0;
                        } else if (all_done && (avail <= 0)) {
                                break;
                        } else {
                                __nanosleep(2000);
                                idle++;
                        }
                }
        } else {
                /* ---- stepping warp ---------------------------------------- */
                WarpPool pool;
                pool.words = cta_words + warp * warp_slot_words(n_slots);
                pool.phase = pools + warp * warp_pool_bytes(n_slots);
                pool.list = pool.phase + n_slots;
                pool.slots = n_slots;
                const unsigned int lt_mask = (1u << lane) - 1u;
                bool exhausted = false;
                bool counted = false;
                unsigned int round = 0;

                for (;; round++) {
                        /* regroup the slots by pending work */
                        int count[N_PHASES];
#pragma unroll
                        for (int ph = 0; ph < N_PHASES; ph++) count[ph] = 0;
                        for (int c = 0; c < n_slots / 32; c++) {
                                const int slot = c * 32 + lane;
                                const int mine =
                                    ((volatile unsigned char *)pool.phase)[slot];
#pragma unroll
                                for (int ph = 0; ph < N_PHASES; ph++) {
                                        const unsigned int m = __ballot_sync(
                                            0xffffffffu, mine == ph);
                                        if (mine == ph)
                                                pool.list[ph * n_slots +
                                                    count[ph] +
                                                    __popc(m & lt_mask)] =
                                                    (unsigned char)slot;
                                        count[ph] += __popc(m);
                                }
                        }
                        __syncwarp();
                        const int n_step = count[PHASE_STEP];
                        const int n_ehs = count[PHASE_EHS];
                        const int n_del = count[PHASE_DEL];
                        const int n_empty = count[PHASE_EMPTY];
                        if ((n_empty == n_slots) && exhausted && !counted) {
                                counted = true;
                                if (lane == 0)
                                        atomicAdd((unsigned int *)&sh.n_done, 1u);
                        }
                        /* lockstep: the stepping warps start their passes
                         * together (and leave together) */
                        if ((round % WF_SYNC_ROUNDS) == 0) {
                                named_barrier_sync(1, 32 * WF_STEP_WARPS);
                                if (sh.n_done == WF_STEP_WARPS) break;
                        }
                        stat[STAT_ROUNDS]++;

                        /* refill the empty slots from the global cursor */
                        if ((n_empty > 0) && !exhausted) {
                                long base = 0;
                                if (lane == 0)
                                        base = (long)atomicAdd(a.cursor,
                                            (unsigned long long)n_empty);
                                base = __shfl_sync(0xffffffffu, base, 0);
                                if (base + n_empty >= a.n) exhausted = true;
                                for (int j = lane; j < n_empty; j += 32) {
                                        const long i = base + j;
                                        if (i >= a.n) continue;
                                        const int slot = pool.list[
                                            PHASE_EMPTY * n_slots + j];
                                        Slot s;
                                        s.index = i;
                                        s.state = a.states[i];
                                        s.rng.init(
                                            a.seed, a.offset + (uint64_t)i);
                                        Outcome out;
                                        const int ph = track_begin(p, cfg, geo,
                                            s.rng, s.state, s.trk, s.lp, out);
                                        s.medium0 = out.medium[0] & 0xff;
                                        if (ph == PHASE_DONE) {
                                                track_flush(s, out, a.states,
                                                    a.events, a.media,
                                                    a.status, a.n_random,
                                                    a.n_steps);
                                        } else {
                                                slot_store(s, pool.words,
                                                    n_slots, slot, a.seed);
                                                pool.phase[slot] =
                                                    (unsigned char)ph;
                                        }
                                }
                                stat[STAT_REFILLS]++;
                        }

                        /* hard inelastic events: hand them to the DEL warps */
                        if ((WF_DEL_WARPS > 0) && (n_del > 0)) {
                                __threadfence();
                                for (int j = lane; j < n_del; j += 32) {
                                        const int slot = pool.list[
                                            PHASE_DEL * n_slots + j];
                                        const unsigned int pos = atomicAdd(
                                            (unsigned int *)&sh.del_tail, 1u);
                                        del_queue[pos % queue_size] =
                                            (unsigned short)((warp << 8) | slot);
                                        pool.phase[slot] =
                                            (unsigned char)PHASE_DELQ;
                                }
                        }

                        /* Monte Carlo steps */
                        if (n_step > 0) {
                                stat[STAT_STEP_PASSES]++;
                                stat[STAT_STEP_LANES] +=
                                    (n_step < 32) ? n_step : 32;
                        }
                        if (lane < n_step) {
                                const int slot =
                                    pool.list[PHASE_STEP * n_slots + lane];
                                Slot s;
                                slot_load(s, pool.words, n_slots, slot, a.seed);
                                const int ph = stepping_step(
                                    p, cfg, geo, s.trk, s.rng, s.state, s.lp);
                                slot_retire(s, ph, pool, slot, a);
                        }

                        /* hard elastic events */
#if PB_WF_SYNC_ROUNDS == 1 && !defined(PB_WF_NO_EHS_BARRIER)
                        named_barrier_sync(3, 32 * WF_STEP_WARPS);
#endif
                        /* Served in the same rounds by every warp (lockstep:
                         * a round lasts as long as its slowest warp), or
                         * when the warp is running out of steps. */
                        if ((n_ehs > 0) &&
                            ((a.ehs_period > 0) ?
                                    (((round % a.ehs_period) ==
                                         (unsigned int)(a.ehs_period - 1)) ||
                                        (n_step < 16)) :
                                    ((n_ehs >= a.ehs_threshold) ||
                                        (n_step < 32)))) {
                                stat[STAT_EHS_PASSES] += (n_ehs + 31) / 32;
                                stat[STAT_EHS_LANES] += n_ehs;
                                for (int j = lane; j < n_ehs; j += 32) {
                                        const int slot = pool.list[
                                            PHASE_EHS * n_slots + j];
                                        Slot s;
                                        slot_load(s, pool.words, n_slots, slot,
                                            a.seed);
                                        const int ph = stepping_ehs(p, cfg, geo,
                                            s.trk, s.rng, s.state, s.lp);
                                        slot_retire(s, ph, pool, slot, a);
                                }
                        }

                        /* without DEL warps: serve the own events */
                        if ((WF_DEL_WARPS == 0) &&
                            ((n_del >= a.del_threshold) ||
                                ((n_del > 0) && (n_step == 0) &&
                                    (n_ehs == 0)))) {
                                stat[STAT_DEL_PASSES] += (n_del + 31) / 32;
                                stat[STAT_DEL_LANES] += n_del;
                                for (int j = lane; j < n_del; j += 32) {
                                        const int slot = pool.list[
                                            PHASE_DEL * n_slots + j];
                                        Slot s;
                                        slot_load(s, pool.words, n_slots, slot,
                                            a.seed);
                                        const int ph = stepping_del(p, cfg, geo,
                                            s.trk, s.rng, s.state, s.lp);
                                        slot_retire(s, ph, pool, slot, a);
                                }
                        }
                        __syncwarp();
                }
        }
        if ((a.stats != nullptr) && (lane == 0)) {
#pragma unroll
                for (int k = 0; k < N_STATS; k++)
                        atomicAdd(a.stats + k, stat[k]);
        }
}

/* Scheduler tunables (environment overrides are for tuning runs). */
struct WavefrontTunables {
        int slots, ehs_threshold, ehs_period, del_threshold;
        unsigned long long * stats;
        int sms;
};

inline const WavefrontTunables & wavefront_tunables()
{
        static WavefrontTunables t = { 0, 0, 0, 0, nullptr, 0 };
        if (t.slots != 0) return t;
        const char * v;
        int slots = WF_SLOTS_DEFAULT;
        t.ehs_threshold = WF_EHS_THRESHOLD;
        t.del_threshold = WF_DEL_THRESHOLD;
        if ((v = getenv("PUMAS_B200_WF_SLOTS")) != nullptr) slots = atoi(v);
        if ((v = getenv("PUMAS_B200_WF_EHS")) != nullptr)
                t.ehs_threshold = atoi(v);
        if ((v = getenv("PUMAS_B200_WF_DEL")) != nullptr)
                t.del_threshold = atoi(v);
        t.ehs_period = WF_EHS_PERIOD;
        if ((v = getenv("PUMAS_B200_WF_EHS_PERIOD")) != nullptr)
                t.ehs_period = atoi(v);
        slots = ((slots + 31) / 32) * 32;
        if (slots < 32) slots = 32;
        if (slots > WF_SLOTS_MAX) slots = WF_SLOTS_MAX;
        if (getenv("PUMAS_B200_WF_STATS") != nullptr)
                cudaMalloc(&t.stats, N_STATS * sizeof *t.stats);
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&t.sms, cudaDevAttrMultiProcessorCount, dev);
        t.slots = slots;
        return t;
}

/* The workspace holds the work cursor and the parked tracks of the largest
 * persistent grid any instantiation may use. */
constexpr int WF_MAX_CTAS_PER_SM = 4;

inline size_t wavefront_workspace_bytes()
{
        const WavefrontTunables & t = wavefront_tunables();
        return 256 +
            (size_t)t.sms * WF_MAX_CTAS_PER_SM * WF_STEP_WARPS *
            warp_slot_words(t.slots) * 8;
}

template <class Cfg, class Geo>
cudaError_t launch_wavefront(const Physics & p, const Cfg & cfg,
    const Geo & geo, uint64_t seed, uint64_t offset, long n, State * states,
    int * events, int * media, int * status, unsigned int * n_random,
    unsigned int * n_steps, void * workspace, cudaStream_t stream)
{
        static int max_grid = 0; /* per instantiation */
        const WavefrontTunables & t = wavefront_tunables();
        const size_t smem = sizeof(WavefrontShared) +
            ((2 * (size_t)WF_STEP_WARPS * t.slots + 15) & ~(size_t)15) +
            warp_pool_bytes(t.slots) * WF_STEP_WARPS;
        cudaError_t e;
        if (max_grid == 0) {
                e = cudaFuncSetAttribute(transport_wavefront_kernel<Cfg, Geo>,
                    cudaFuncAttributePreferredSharedMemoryCarveout,
                    16 /* percent of the 228 KB: work lists only */);
                if (e != cudaSuccess) return e;
                int per_sm = 0;
                e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm,
                    transport_wavefront_kernel<Cfg, Geo>, WF_THREADS, smem);
                if (e != cudaSuccess) return e;
                if (per_sm < 1) per_sm = 1;
                if (per_sm > WF_MAX_CTAS_PER_SM) per_sm = WF_MAX_CTAS_PER_SM;
                max_grid = t.sms * per_sm;
        }
        /* Persistent grid: one wave of resident CTAs, fewer for small n. */
        const long per_cta = (long)t.slots * WF_STEP_WARPS;
        const long want = (n + per_cta - 1) / per_cta;
        const int grid = (want < max_grid) ? (int)want : max_grid;
        unsigned long long * cursor =
            static_cast<unsigned long long *>(workspace);
        e = cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return e;
        if (t.stats != nullptr)
                cudaMemsetAsync(t.stats, 0, N_STATS * sizeof *t.stats, stream);
        WavefrontArgs a;
        a.seed = seed;
        a.offset = offset;
        a.n = n;
        a.states = states;
        a.events = events;
        a.media = media;
        a.status = status;
        a.n_random = n_random;
        a.n_steps = n_steps;
        a.cursor = cursor;
        a.slot_words = cursor + 32;
        a.stats = t.stats;
        a.slots = t.slots;
        a.ehs_threshold = t.ehs_threshold;
        a.ehs_period = t.ehs_period;
        a.del_threshold = t.del_threshold;
        transport_wavefront_kernel<Cfg, Geo>
            <<<grid, WF_THREADS, smem, stream>>>(p, cfg, geo, a);
        e = cudaGetLastError();
        if ((e == cudaSuccess) && (t.stats != nullptr)) {
                unsigned long long h[N_STATS];
                cudaMemcpyAsync(h, t.stats, sizeof h, cudaMemcpyDeviceToHost,
                    stream);
                cudaStreamSynchronize(stream);
                fprintf(stderr,
                    "[wavefront] grid=%d slots=%d ehs_thr=%d del_thr=%d "
                    "rounds=%llu step: passes=%llu lanes/pass=%.2f ehs: "
                    "passes=%llu lanes/pass=%.2f del: passes=%llu "
                    "lanes/pass=%.2f refills=%llu\n",
                    grid, t.slots, t.ehs_threshold, t.del_threshold,
                    h[STAT_ROUNDS], h[STAT_STEP_PASSES],
                    (double)h[STAT_STEP_LANES] /
                        (double)(h[STAT_STEP_PASSES] ? h[STAT_STEP_PASSES] : 1),
                    h[STAT_EHS_PASSES],
                    (double)h[STAT_EHS_LANES] /
                        (double)(h[STAT_EHS_PASSES] ? h[STAT_EHS_PASSES] : 1),
                    h[STAT_DEL_PASSES],
                    (double)h[STAT_DEL_LANES] /
                        (double)(h[STAT_DEL_PASSES] ? h[STAT_DEL_PASSES] : 1),
                    h[STAT_REFILLS]);
        }
        return e;
}

/* Specialised instantiations (kernels_spec*.cu), tried in turn by
 * launch_transport before the general kernel. Each returns false when the
 * settings / geometry are not its own. */
#define PB_SPECIALISATION(NAME)                                                \
        bool NAME(const Physics & p, const Settings & cfg,                     \
            const Geometry & geo, uint64_t seed, uint64_t offset, long n,      \
            State * states, int * events, int * media, int * status,           \
            unsigned int * n_random, unsigned int * n_steps, void * workspace, \
            cudaStream_t stream, cudaError_t * error)

/* Is the geometry a single infinite medium without locals callback? */
inline bool geometry_is_plain_uniform(const Geometry & geo)
{
        return (geo.kind == PUMAS_B200_GEOMETRY_UNIFORM) &&
            !geo.layer[0].has_locals;
}

#define PB_DEFINE_SPECIALISATION(NAME, EL, SC, DI, DE)                         \
        PB_SPECIALISATION(NAME)                                                \
        {                                                                      \
                typedef StaticSettings<EL, SC, DI, DE> Cfg;                    \
                if (!Cfg::matches(cfg) || !geometry_is_plain_uniform(geo))     \
                        return false;                                          \
                UniformGeometry ug;                                            \
                ug.material = geo.layer[0].material;                           \
                ug.step_type = geo.step_type;                                  \
                *error = launch_wavefront(p, Cfg(cfg), ug, seed, offset, n,    \
                    states, events, media, status, n_random, n_steps,          \
                    workspace, stream);                                        \
                return true;                                                   \
        }

} /* namespace pb */

#endif
