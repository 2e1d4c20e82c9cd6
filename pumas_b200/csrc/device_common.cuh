/*
 * device_common.cuh -- helpers shared by the transport kernels: the Philox
 * block function and the register resident stream of the monolithic kernel,
 * bit packing, named barriers, the batch arguments and the tuning knobs.
 */
#ifndef PB_DEVICE_COMMON_CUH
#define PB_DEVICE_COMMON_CUH

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "kernels.h"
#include "pb_transport.h"

namespace pb {

/* One Philox4x32-10 block. */
static __device__ __forceinline__ uint4 philox_block_inline(
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

/* The same out of line (one copy instead of one per call site of uniform()). */
static __device__ __noinline__ uint4 philox_block(
    uint32_t key0, uint32_t key1, uint32_t block, uint32_t trk0, uint32_t trk1)
{
        return philox_block_inline(key0, key1, block, trk0, trk1);
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

__device__ __forceinline__ void named_barrier_sync(int id, int threads)
{
        asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

/* One batch of tracks: what every kernel receives. */
struct BatchArgs {
        uint64_t seed, offset;
        long n;
        State * states;
        int * events;
        int * media;
        int * status;
        unsigned int * n_random;
        unsigned int * n_steps;
        unsigned long long * cursor; /* global work cursor (workspace) */
        unsigned char * workspace;   /* WORKSPACE_BYTES of global scratch */
        int n_del_ctas; /* CTAs serving the hard inelastic events (offload) */
        unsigned long long * stats;  /* optional scheduler statistics */
        int ehs_period;    /* the warp pools are served every that many rounds */
        int del_threshold; /* events queued before a DEL warp starts a pass */
        int service_patience; /* polls (2 us each) a service warp waits for
                               * that many before it serves what there is */
};

constexpr int N_STATS = 16;

/* Scheduler knobs. The environment overrides are for tuning runs; they are
 * read once per process. */
struct Tunables {
        int ehs_period, del_threshold; /* <= 0: the kernel's default */
        int grid_cap;                  /* <= 0: one CTA per SM */
        int del_ctas;                  /* <= 0: the launcher's choice */
        int service_patience;          /* <= 0: the kernel's default */
        unsigned long long * stats;
        int sms;
};

inline const Tunables & tunables()
{
        static Tunables t = { 0, 0, 0, 0, 0, nullptr, 0 };
        static std::once_flag once;
        std::call_once(once, []() {
        const char * v;
        if ((v = getenv("PUMAS_B200_EHS_PERIOD")) != nullptr)
                t.ehs_period = atoi(v);
        if ((v = getenv("PUMAS_B200_DEL_THRESHOLD")) != nullptr)
                t.del_threshold = atoi(v);
        if ((v = getenv("PUMAS_B200_GRID")) != nullptr) t.grid_cap = atoi(v);
        if ((v = getenv("PUMAS_B200_DEL_CTAS")) != nullptr)
                t.del_ctas = atoi(v);
        if ((v = getenv("PUMAS_B200_SERVICE_PATIENCE")) != nullptr)
                t.service_patience = atoi(v);
        if (getenv("PUMAS_B200_STATS") != nullptr)
                cudaMalloc(&t.stats, N_STATS * sizeof *t.stats);
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&t.sms, cudaDevAttrMultiProcessorCount, dev);
        });
        return t;
}

/* Per device facts (a host process may drive several GPUs). */
constexpr int MAX_DEVICES = 64;
inline int current_device()
{
        int dev = 0;
        cudaGetDevice(&dev);
        return ((dev >= 0) && (dev < MAX_DEVICES)) ? dev : 0;
}
inline int device_sms(int dev)
{
        static int sms[MAX_DEVICES] = { 0 };
        if (sms[dev] == 0)
                cudaDeviceGetAttribute(
                    &sms[dev], cudaDevAttrMultiProcessorCount, dev);
        return sms[dev];
}

/* Kernels launched by the last launch_transport of this process (diagnostics:
 * pumas_b200_context_last_kernel_time). */
inline int & last_launch_count()
{
        /* (per host thread: one context per thread, as the reference) */
        static thread_local int n = 0;
        return n;
}

/* The workspace holds the global work cursor and, for the kernels that hand
 * the hard inelastic events over to dedicated CTAs (resident.cuh,
 * PB_RS_OFFLOAD; hot_kernel.cuh), their queues and parking area / the mirrors
 * of the slots. */
constexpr size_t WORKSPACE_BYTES = (size_t)32 << 20;
inline size_t workspace_bytes() { return WORKSPACE_BYTES; }

} /* namespace pb */

#endif
