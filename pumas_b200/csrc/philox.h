/*
 * philox.h -- counter based Philox4x32-10 stream, one per track.
 *
 * Plain C99 / C++ / CUDA. Replaces the per-context Mersenne Twister of the
 * reference (src/pumas.c:2330-2542) behind `context->random` (pumas.h:551).
 * The same file is compiled into the CPU oracle driver so both sides draw the
 * identical sequence.
 *
 * Stream definition (the contract tests rely on):
 *   key     = (seed & 0xffffffff, seed >> 32)
 *   counter = (block, 0, track & 0xffffffff, track >> 32), block = 0, 1, 2...
 *   the four 32 bit words of a block are consumed in order w0, w1, w2, w3
 *   u = (w + 0.5) * 2^-32, strictly inside (0, 1) and exact in binary64.
 * The reference maps a 32 bit MT word to [0, 1] (pumas.c:2541); the open
 * interval removes the log(0) hazards at pumas.c:6608, 6639, 6919.
 *
 * Algorithm: Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
 * SC'11 (Random123); known answer vectors are pinned in tests/test_philox.py.
 */
#ifndef PUMAS_B200_PHILOX_H
#define PUMAS_B200_PHILOX_H

#include <stdint.h>

#if defined(__CUDACC__)
#define PB_HD __host__ __device__ __forceinline__
#else
#define PB_HD static inline
#endif

struct pb_philox {
        uint32_t key[2];
        uint32_t track[2];
        uint32_t block;   /* index of the next block to generate */
        uint32_t have;    /* unread words left in out[] */
        uint32_t out[4];
        uint32_t count;   /* variates drawn so far (diagnostics) */
};

PB_HD void pb_philox_mulhilo(
    uint32_t a, uint32_t b, uint32_t * hi, uint32_t * lo)
{
#if defined(__CUDA_ARCH__)
        *lo = a * b;
        *hi = __umulhi(a, b);
#else
        const uint64_t p = (uint64_t)a * (uint64_t)b;
        *lo = (uint32_t)p;
        *hi = (uint32_t)(p >> 32);
#endif
}

/* One Philox4x32-10 block: ctr, key -> out. */
PB_HD void pb_philox4x32_10(
    const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
        uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
        uint32_t k0 = key[0], k1 = key[1];
        int r;
        for (r = 0; r < 10; r++) {
                uint32_t hi0, lo0, hi1, lo1;
                pb_philox_mulhilo(0xD2511F53u, c0, &hi0, &lo0);
                pb_philox_mulhilo(0xCD9E8D57u, c2, &hi1, &lo1);
                c0 = hi1 ^ c1 ^ k0;
                c1 = lo1;
                c2 = hi0 ^ c3 ^ k1;
                c3 = lo0;
                k0 += 0x9E3779B9u;
                k1 += 0xBB67AE85u;
        }
        out[0] = c0;
        out[1] = c1;
        out[2] = c2;
        out[3] = c3;
}

PB_HD void pb_philox_init(
    struct pb_philox * g, uint64_t seed, uint64_t track)
{
        g->key[0] = (uint32_t)seed;
        g->key[1] = (uint32_t)(seed >> 32);
        g->track[0] = (uint32_t)track;
        g->track[1] = (uint32_t)(track >> 32);
        g->block = 0;
        g->have = 0;
        g->count = 0;
        g->out[0] = g->out[1] = g->out[2] = g->out[3] = 0;
}

/* Next uniform variate in the open interval (0, 1). */
PB_HD double pb_philox_uniform(struct pb_philox * g)
{
        if (g->have == 0) {
                uint32_t ctr[4];
                ctr[0] = g->block;
                ctr[1] = 0;
                ctr[2] = g->track[0];
                ctr[3] = g->track[1];
                pb_philox4x32_10(ctr, g->key, g->out);
                g->block++;
                g->have = 4;
        }
        const uint32_t w = g->out[4 - g->have];
        g->have--;
        g->count++;
        return ((double)w + 0.5) * (1.0 / 4294967296.0);
}

#endif
