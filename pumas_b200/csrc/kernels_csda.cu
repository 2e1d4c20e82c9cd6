/*
 * kernels_csda.cu -- the closed form CSDA kernels (BASELINE.json configs[1]:
 * "forward CSDA-only transport ... pure table-lookup path").
 *
 * With energy loss = CSDA, no scattering and a single uniform medium without
 * locals, every track ends in transport_with_csda (pumas.c:4777-4949): half a
 * dozen PCHIP look-ups, no loop, no random number. The path is HBM bound:
 * 104 B of state in, 104 B out, 4 B of event per track (+ 16 B of optional
 * media and counters).
 *
 * transport_csda_tile_kernel (product path): persistent CTAs, one per SM. The
 * states move HBM -> shared memory -> HBM as TILES of CSDA_TILE records
 * (53 KB) with the bulk asynchronous copy engine (cp.async.bulk + mbarrier;
 * UBLKCP in SASS), three buffers deep: while the threads work on tile k from
 * shared memory, tile k + 1 is in flight in and tile k - 1 in flight out, so
 * the table look-ups of a tile overlap the copies of its neighbours. A thread
 * reads its 104 B record as 13 conflict free 8 byte words (stride 13 words,
 * odd). The look-ups themselves are pb_csda.h: O(1) locators and per cell
 * packed PCHIP coefficients staged in shared memory (37 KB), instead of six
 * bisections of eight dependent loads each (the long scoreboard was the first
 * stall of the round 1 kernel, profiles/r1_csda_ncu.txt).
 *
 * transport_csda_kernel: one thread per track on the general code; serves the
 * states that are not 16 byte aligned and the tail of a batch, and the A/B
 * comparison (same results bit for bit).
 */
#define PB_HOT_SMEM_TABLES
#include "resident.cuh"
#include "pb_csda.h"

namespace pb {

constexpr int CSDA_BLOCK = 256;
#ifndef PB_CSDA_MIN_CTAS
#define PB_CSDA_MIN_CTAS 4
#endif

template <class Cfg>
__global__ void __launch_bounds__(CSDA_BLOCK, PB_CSDA_MIN_CTAS)
    transport_csda_kernel(const Physics p, const Cfg cfg,
        const UniformGeometry geo, const BatchArgs a, long first)
{
        const long i = first + (long)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= a.n) return;

        State state = a.states[i];
        DeviceRng rng;
        rng.init(a.seed, a.offset + (uint64_t)i);
        unsigned int steps = 0;
        const Outcome out = transport_track(p, cfg, geo, rng, state, &steps);
        a.states[i] = state;
        if (a.events != nullptr) a.events[i] = out.event;
        if (a.media != nullptr) {
                a.media[2 * i] = out.medium[0];
                a.media[2 * i + 1] = out.medium[1];
        }
        if (out.rc != RETURN_SUCCESS) atomicCAS(a.status, 0, out.rc);
        if (a.n_random != nullptr) a.n_random[i] = rng.count;
        if (a.n_steps != nullptr) a.n_steps[i] = steps;
}

/* ---- tiles through the bulk copy engine ------------------------------------ */
constexpr int CSDA_TILE = 512;    /* tracks per tile = threads per CTA */
constexpr int CSDA_BUFFERS = 3;
constexpr int CSDA_TILE_BYTES = CSDA_TILE * (int)sizeof(State);
static_assert(sizeof(State) == 104, "struct pumas_state layout");
static_assert(CSDA_TILE_BYTES % 16 == 0, "bulk copies move 16 byte units");

/* shared memory, in 64 bit words: [mbarriers][tiles][tables] */
constexpr unsigned int CSDA_TILES0 = 8;
constexpr unsigned int CSDA_TABLES0 =
    CSDA_TILES0 + CSDA_BUFFERS * CSDA_TILE_BYTES / 8;

inline size_t csda_smem_bytes(int n_cells)
{
        return (size_t)CSDA_TABLES0 * 8 +
            (size_t)n_cells * (CSDAK_WORDS + CSDAX_WORDS) * 8;
}

__device__ __forceinline__ uint32_t smem_address(const void * p)
{
        return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, int count)
{
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar),
                     "r"(count)
                     : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
        asm volatile(
            "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
            "r"(bytes)
            : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
        asm volatile("{\n"
                     ".reg .pred P1;\n"
                     "CSDA_WAIT:\n"
                     "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                     "@P1 bra CSDA_DONE;\n"
                     "bra CSDA_WAIT;\n"
                     "CSDA_DONE:\n"
                     "}" ::"r"(bar),
                     "r"(parity)
                     : "memory");
}

/* global -> shared, completion on the mbarrier */
__device__ __forceinline__ void bulk_load(
    uint32_t dst, const void * src, uint32_t bytes, uint32_t bar)
{
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::"
                     "bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(src), "r"(bytes), "r"(bar)
                     : "memory");
}

/* shared -> global, completion through the bulk async-group */
__device__ __forceinline__ void bulk_store(
    void * dst, uint32_t src, uint32_t bytes)
{
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::
                         "l"(dst),
                     "r"(src), "r"(bytes)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template <class Cfg>
__global__ void __launch_bounds__(CSDA_TILE, 1)
    transport_csda_tile_kernel(const __grid_constant__ Physics p,
        const __grid_constant__ Cfg cfg,
        const __grid_constant__ UniformGeometry geo,
        const __grid_constant__ CsdaCtx c, const __grid_constant__ BatchArgs a,
        long n_tiles)
{
        unsigned char * const smem_raw =
            reinterpret_cast<unsigned char *>(hot_smem64);
        const uint32_t bars = smem_address(smem_raw);
        const uint32_t tiles = smem_address(smem_raw + CSDA_TILES0 * 8);

        if (threadIdx.x == 0) {
                for (int b = 0; b < CSDA_BUFFERS; b++)
                        mbar_init(bars + 8u * b, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        /* stage the packed tables of the medium's material */
        for (int k = threadIdx.x; k < c.n_cells * CSDAK_WORDS; k += CSDA_TILE)
                hot_smem64[c.cK + k] =
                    (unsigned long long)__double_as_longlong(__ldg(c.gK + k));
        for (int k = threadIdx.x; k < c.n_cells * CSDAX_WORDS; k += CSDA_TILE)
                hot_smem64[c.cX + k] =
                    (unsigned long long)__double_as_longlong(__ldg(c.gX + k));
        __syncthreads();

        /* tiles of this CTA: blockIdx.x, + gridDim.x, ... */
        const long first = blockIdx.x;
        const long stride = gridDim.x;
        if ((threadIdx.x == 0) && (first < n_tiles)) {
                mbar_expect_tx(bars, CSDA_TILE_BYTES);
                bulk_load(tiles, a.states + first * CSDA_TILE, CSDA_TILE_BYTES,
                    bars);
        }
        long k = 0;
        for (long tile = first; tile < n_tiles; tile += stride, k++) {
                const int buffer = (int)(k % CSDA_BUFFERS);
                const uint32_t parity = (uint32_t)((k / CSDA_BUFFERS) & 1);
                /* the next tile, into the buffer tile k - 2 was stored from:
                 * that store must have read it */
                if ((threadIdx.x == 0) && (tile + stride < n_tiles)) {
                        const int next = (int)((k + 1) % CSDA_BUFFERS);
                        asm volatile("cp.async.bulk.wait_group.read 1;" ::
                                         : "memory");
                        mbar_expect_tx(bars + 8u * next, CSDA_TILE_BYTES);
                        bulk_load(tiles + (uint32_t)next * CSDA_TILE_BYTES,
                            a.states + (tile + stride) * CSDA_TILE,
                            CSDA_TILE_BYTES, bars + 8u * next);
                }
                mbar_wait(bars + 8u * buffer, parity);

                /* this thread's record: 13 words at a stride of 13 */
                unsigned long long * const record = hot_smem64 + CSDA_TILES0 +
                    (size_t)buffer * (CSDA_TILE_BYTES / 8) +
                    (size_t)threadIdx.x * 13;
                State state;
                {
                        unsigned long long w[13];
#pragma unroll
                        for (int q = 0; q < 13; q++) w[q] = record[q];
                        memcpy(&state, w, sizeof state);
                }
                const long i = tile * CSDA_TILE + threadIdx.x;
                Outcome out;
                unsigned int steps = 0, drawn = 0;
                if (!csda_track(cfg, c, state, out)) {
                        DeviceRng rng;
                        rng.init(a.seed, a.offset + (uint64_t)i);
                        out = transport_track(p, cfg, geo, rng, state, &steps);
                        drawn = rng.count;
                }
                {
                        unsigned long long w[13];
                        memcpy(w, &state, sizeof state);
#pragma unroll
                        for (int q = 0; q < 13; q++) record[q] = w[q];
                }
                if (a.events != nullptr) a.events[i] = out.event;
                if (a.media != nullptr) {
                        a.media[2 * i] = out.medium[0];
                        a.media[2 * i + 1] = out.medium[1];
                }
                if (out.rc != RETURN_SUCCESS) atomicCAS(a.status, 0, out.rc);
                if (a.n_random != nullptr) a.n_random[i] = drawn;
                if (a.n_steps != nullptr) a.n_steps[i] = steps;

                /* the tile goes back: the writes of the threads must be
                 * visible to the copy engine */
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
                if (threadIdx.x == 0)
                        bulk_store(a.states + tile * CSDA_TILE,
                            tiles + (uint32_t)buffer * CSDA_TILE_BYTES,
                            CSDA_TILE_BYTES);
        }
        if (threadIdx.x == 0)
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <class Cfg>
static cudaError_t launch_csda(const Physics & p, const Physics & host_view,
    const Cfg & cfg, const UniformGeometry & ug, const BatchArgs & a,
    int kernel, cudaStream_t stream)
{
        long done = 0;
        const int n_cells = host_view.n_energies - 1;
        const size_t smem = csda_smem_bytes(n_cells);
        const long n_tiles = a.n / CSDA_TILE;
        if ((kernel == KERNEL_RESIDENT) && (n_tiles > 0) &&
            (smem <= (size_t)227 * 1024) &&
            ((reinterpret_cast<uintptr_t>(a.states) & 15) == 0)) {
                cudaError_t e = cudaFuncSetAttribute(
                    transport_csda_tile_kernel<Cfg>,
                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
                CsdaCtx c = csda_context(host_view, p, cfg, ug.material, ug.density);
                c.cK = CSDA_TABLES0;
                c.cX = CSDA_TABLES0 + (unsigned int)(n_cells * CSDAK_WORDS);
                const int sms = device_sms(current_device());
                const int grid = (n_tiles < sms) ? (int)n_tiles : sms;
                transport_csda_tile_kernel<Cfg><<<grid, CSDA_TILE, smem,
                    stream>>>(p, cfg, ug, c, a, n_tiles);
                e = cudaGetLastError();
                if (e != cudaSuccess) return e;
                done = n_tiles * CSDA_TILE;
                last_launch_count() = 1;
        }
        if (done < a.n) {
                /* the tail of the batch (or all of it) */
                const long blocks = (a.n - done + CSDA_BLOCK - 1) / CSDA_BLOCK;
                transport_csda_kernel<Cfg><<<(unsigned int)blocks, CSDA_BLOCK, 0,
                    stream>>>(p, cfg, ug, a, done);
                last_launch_count() = (done > 0) ? 2 : 1;
        }
        return cudaGetLastError();
}

/* Closed form CSDA in a plain uniform medium, decays weighted (the default).
 * Returns false when the settings / geometry are not of that kind. */
bool launch_spec_csda(const Physics & p, const Physics & host_view,
    const Settings & cfg, const Geometry & geo, uint64_t seed, uint64_t offset,
    long n, State * states, int * events, int * media, int * status,
    unsigned int * n_random, unsigned int * n_steps, void * workspace,
    int kernel, cudaStream_t stream, cudaError_t * error)
{
        (void)workspace;
        if ((cfg.energy_loss != MODE_CSDA) ||
            (cfg.scattering != MODE_DISABLED) ||
            (cfg.decay != MODE_WEIGHTED))
                return false;
        UniformGeometry ug;
        if (!uniform_geometry_of(geo, &ug)) return false;
        BatchArgs a;
        a.seed = seed;
        a.offset = offset;
        a.n = n;
        a.states = states;
        a.events = events;
        a.media = media;
        a.status = status;
        a.n_random = n_random;
        a.n_steps = n_steps;
        a.cursor = nullptr;
        a.stats = nullptr;
        a.ehs_period = 0;
        a.del_threshold = 0;
        if (cfg.direction == MODE_FORWARD) {
                typedef StaticSettings<MODE_CSDA, MODE_DISABLED, MODE_FORWARD,
                    MODE_WEIGHTED>
                    Cfg;
                *error = launch_csda(p, host_view, Cfg(cfg), ug, a, kernel, stream);
        } else {
                typedef StaticSettings<MODE_CSDA, MODE_DISABLED, MODE_BACKWARD,
                    MODE_WEIGHTED>
                    Cfg;
                *error = launch_csda(p, host_view, Cfg(cfg), ug, a, kernel, stream);
        }
        return true;
}

} /* namespace pb */
