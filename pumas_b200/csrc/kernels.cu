/*
 * kernels.cu -- sm_100a transport kernels.
 *
 * One thread owns one track and runs its whole step loop
 * (pb::transport_track, the restatement of pumas_context_transport,
 * src/pumas.c:3114). The physics tables stay in HBM / L2 (1.6 MB, resident in
 * the 126 MB L2 after first touch) and are read through the read-only path;
 * track states are read and written once (104 B each way).
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

        __device__ __forceinline__ void init(uint64_t seed, uint64_t track)
        {
                key0 = (uint32_t)seed;
                key1 = (uint32_t)(seed >> 32);
                trk0 = (uint32_t)track;
                trk1 = (uint32_t)(track >> 32);
                block = have = count = 0;
                o0 = o1 = o2 = o3 = 0;
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

cudaError_t launch_transport(const Physics & p, const Settings & cfg,
    const Geometry & geo, uint64_t seed, uint64_t offset, long n,
    State * states, int * events, int * media, int * status,
    unsigned int * n_random, unsigned int * n_steps, cudaStream_t stream)
{
        if (n <= 0) return cudaSuccess;
        const long blocks = (n + TRANSPORT_BLOCK - 1) / TRANSPORT_BLOCK;
        transport_kernel<<<(unsigned int)blocks, TRANSPORT_BLOCK, 0, stream>>>(
            p, cfg, geo, seed, offset, n, states, events, media, status,
            n_random, n_steps);
        return cudaGetLastError();
}

} /* namespace pb */
