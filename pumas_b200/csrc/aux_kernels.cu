/*
 * aux_kernels.cu -- tallies of a transported batch and the FP64 peak probe.
 *
 * pumas_b200_tally: what a user loop does after pumas_context_transport
 * (examples/pumas/straight.c:175-178: w += wi, w2 += wi * wi), for a whole
 * batch on the device: event counts, sum of weights and of squared weights of
 * the tracks selected by an event mask, and a histogram (counts) of their
 * kinetic energy in log-spaced bins. Sums are reduced in a fixed order (per
 * thread strided partial sums, shared memory tree, then the blocks in index
 * order), so a given (n, grid) gives bit identical results from run to run;
 * these few hundred bytes are what the ranks all-reduce over NCCL.
 *
 * pumas_b200_fp64_peak: dependent-free DFMA chains, the denominator of the
 * roofline (the FP64 pipe bounds this workload).
 */
#include <cuda_runtime.h>

#include <cmath>

#include "pb_common.h"
#include "pb_hot_math.h"
#include "pumas_b200.h"

namespace {

constexpr int TALLY_THREADS = 256;
constexpr int TALLY_BLOCKS = 592;

constexpr int TALLY_SHARED_BINS = 2048;

/* Counts are integers (any order gives the same totals): a thread counts in
 * registers, a CTA in shared memory (histogram of up to TALLY_SHARED_BINS
 * bins), and only the CTA totals go to the global counters. */
__global__ void __launch_bounds__(TALLY_THREADS)
    tally_kernel(const pb::State * __restrict__ states,
        const int * __restrict__ events, long n, int event_mask,
        double ln_emin, double ln_emax, int n_bins,
        double * __restrict__ partial, /* [blocks][2] */
        unsigned long long * __restrict__ counts /* [2 + n_bins] */)
{
        __shared__ double s_w[TALLY_THREADS], s_w2[TALLY_THREADS];
        __shared__ unsigned int s_hist[TALLY_SHARED_BINS];
        __shared__ unsigned long long s_n[2];
        const bool shared_hist = (n_bins <= TALLY_SHARED_BINS);
        if (shared_hist)
                for (int b = threadIdx.x; b < n_bins; b += TALLY_THREADS)
                        s_hist[b] = 0u;
        if (threadIdx.x < 2) s_n[threadIdx.x] = 0ull;
        __syncthreads();

        double w = 0., w2 = 0.;
        unsigned long long n_all = 0ull, n_selected = 0ull;
        const long stride = (long)gridDim.x * blockDim.x;
        for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
             i += stride) {
                const int e = (events != nullptr) ? events[i] : 0;
                const bool selected = (event_mask == 0) || (e & event_mask);
                n_all++;
                if (!selected) continue;
                n_selected++;
                const double wi = states[i].weight;
                w += wi;
                w2 += wi * wi;
                const double k = states[i].energy;
                if ((n_bins > 0) && (k > 0.)) {
                        const double x =
                            (log(k) - ln_emin) / (ln_emax - ln_emin);
                        int b = (int)floor(x * n_bins);
                        if ((b >= 0) && (b < n_bins)) {
                                if (shared_hist)
                                        atomicAdd(&s_hist[b], 1u);
                                else
                                        atomicAdd(&counts[2 + b], 1ull);
                        }
                }
        }
        s_w[threadIdx.x] = w;
        s_w2[threadIdx.x] = w2;
        atomicAdd(&s_n[0], n_all);
        atomicAdd(&s_n[1], n_selected);
        __syncthreads();
        for (int k = TALLY_THREADS / 2; k > 0; k >>= 1) {
                if (threadIdx.x < k) {
                        s_w[threadIdx.x] += s_w[threadIdx.x + k];
                        s_w2[threadIdx.x] += s_w2[threadIdx.x + k];
                }
                __syncthreads();
        }
        if (threadIdx.x == 0) {
                partial[2 * blockIdx.x] = s_w[0];
                partial[2 * blockIdx.x + 1] = s_w2[0];
        }
        if ((threadIdx.x < 2) && (s_n[threadIdx.x] != 0ull))
                atomicAdd(&counts[threadIdx.x], s_n[threadIdx.x]);
        if (shared_hist)
                for (int b = threadIdx.x; b < n_bins; b += TALLY_THREADS)
                        if (s_hist[b] != 0u)
                                atomicAdd(&counts[2 + b],
                                    (unsigned long long)s_hist[b]);
}

__global__ void tally_finish_kernel(const double * __restrict__ partial,
    int blocks, const unsigned long long * __restrict__ counts, int n_counts,
    double * __restrict__ out)
{
        if ((blockIdx.x != 0) || (threadIdx.x != 0)) return;
        double w = 0., w2 = 0.;
        for (int b = 0; b < blocks; b++) {
                w += partial[2 * b];
                w2 += partial[2 * b + 1];
        }
        out[0] = w;
        out[1] = w2;
        for (int k = 0; k < n_counts; k++) out[2 + k] = (double)counts[k];
}

__global__ void fp64_peak_kernel(double * out, int iters, double a, double b)
{
        double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
        double x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
        for (int i = 0; i < iters; i++) {
                x0 = __fma_rn(x0, a, b);
                x1 = __fma_rn(x1, a, b);
                x2 = __fma_rn(x2, a, b);
                x3 = __fma_rn(x3, a, b);
                x4 = __fma_rn(x4, a, b);
                x5 = __fma_rn(x5, a, b);
                x6 = __fma_rn(x6, a, b);
                x7 = __fma_rn(x7, a, b);
        }
        const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
        if (s == 12345.678) out[0] = s; /* keeps the chains alive */
}

} /* namespace */

extern "C" {

/* out: [sum w, sum w^2, n tracks, n selected, counts per bin ...] as doubles
 * (device memory, 4 + n_bins values). scratch: device memory, at least
 * pumas_b200_tally_scratch_bytes(n_bins) bytes. */
long pumas_b200_tally_scratch_bytes(int n_bins)
{
        return (long)(TALLY_BLOCKS * 2 * sizeof(double) +
            (2 + (n_bins > 0 ? n_bins : 0)) * sizeof(unsigned long long));
}

enum pumas_return pumas_b200_tally(const struct pumas_state * states,
    const int * events, long n, int event_mask, double energy_min,
    double energy_max, int n_bins, double * out, void * scratch,
    void * cuda_stream)
{
        if ((states == nullptr) || (out == nullptr) || (scratch == nullptr) ||
            (n < 0) || (n_bins < 0) ||
            ((n_bins > 0) && !(energy_max > energy_min && energy_min > 0.)))
                return PUMAS_RETURN_VALUE_ERROR;
        cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
        double * partial = static_cast<double *>(scratch);
        unsigned long long * counts = reinterpret_cast<unsigned long long *>(
            partial + 2 * TALLY_BLOCKS);
        const int n_counts = 2 + n_bins;
        cudaError_t e = cudaMemsetAsync(
            counts, 0, n_counts * sizeof(unsigned long long), stream);
        if (e != cudaSuccess) return PUMAS_RETURN_INTERNAL_ERROR;
        tally_kernel<<<TALLY_BLOCKS, TALLY_THREADS, 0, stream>>>(
            reinterpret_cast<const pb::State *>(states), events, n, event_mask,
            (n_bins > 0) ? log(energy_min) : 0.,
            (n_bins > 0) ? log(energy_max) : 1., n_bins, partial, counts);
        tally_finish_kernel<<<1, 32, 0, stream>>>(
            partial, TALLY_BLOCKS, counts, n_counts, out);
        return (cudaGetLastError() == cudaSuccess) ?
            PUMAS_RETURN_SUCCESS :
            PUMAS_RETURN_INTERNAL_ERROR;
}

/* Self test of the division and square root of the lean step (pb_hot_math.h:
 * the fast paths of the IEEE routines without their range checks) against
 * a / b and sqrt(x) on random operands: mantissas uniform, exponents uniform
 * over +-`spread` binades (quotients stay inside the stated domain for spread
 * <= 480), one operand in 64 a zero numerator. mismatches[0]: divisions,
 * [1]: square roots whose bits differ. */
namespace {
__device__ __forceinline__ unsigned long long splitmix(unsigned long long & s)
{
        unsigned long long z = (s += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
}

__device__ __forceinline__ double random_operand(
    unsigned long long & s, int spread)
{
        unsigned long long m = splitmix(s);
        const unsigned long long k = splitmix(s);
        const int e = (int)((k >> 8) % (unsigned long long)(2 * spread + 1)) -
            spread;
        /* one operand in 4 next to a power of two (mantissa 0 + j or all ones
         * - j, j < 64): where the ties of the roundings live */
        if ((k & 3ull) == 0)
                m = (k & 4ull) ? ((k >> 3) & 63ull) :
                                 (0x000fffffffffffffull - ((k >> 3) & 63ull));
        const unsigned long long bits = (m & 0x800fffffffffffffull) |
            ((unsigned long long)(1023 + e) << 52);
        return __longlong_as_double((long long)bits);
}

__global__ void hot_math_selftest_kernel(long per_thread,
    unsigned long long seed, int spread, unsigned long long * mismatches)
{
        unsigned long long s = seed +
            0x632be59bd9b4e019ull *
                (unsigned long long)(blockIdx.x * blockDim.x + threadIdx.x + 1);
        unsigned long long bad_div = 0, bad_sqrt = 0;
        for (long i = 0; i < per_thread; i++) {
                double a = random_operand(s, spread);
                const double b = random_operand(s, spread);
                if ((splitmix(s) & 63ull) == 0) a = 0.;
                const double q = pb::hot_div(a, b), q0 = a / b;
                if ((__double_as_longlong(q) != __double_as_longlong(q0)) &&
                    !((q == 0.) && (q0 == 0.)))
                        bad_div++;
                const double x = fabs(random_operand(s, 2 * spread));
                if (__double_as_longlong(pb::hot_sqrt(x)) !=
                    __double_as_longlong(sqrt(x)))
                        bad_sqrt++;
        }
        if (bad_div) atomicAdd(&mismatches[0], bad_div);
        if (bad_sqrt) atomicAdd(&mismatches[1], bad_sqrt);
}
} /* namespace */

enum pumas_return pumas_b200_hot_math_selftest(
    long n, unsigned long seed, int spread, unsigned long long * mismatches)
{
        if ((mismatches == nullptr) || (n < 0) || (spread < 0) ||
            (spread > 480))
                return PUMAS_RETURN_VALUE_ERROR;
        unsigned long long * d = nullptr;
        if (cudaMalloc(&d, 2 * sizeof *d) != cudaSuccess)
                return PUMAS_RETURN_INTERNAL_ERROR;
        cudaMemset(d, 0, 2 * sizeof *d);
        const int threads = 256, blocks = 592;
        const long per_thread = (n + (long)threads * blocks - 1) /
            ((long)threads * blocks);
        hot_math_selftest_kernel<<<blocks, threads>>>(
            per_thread, (unsigned long long)seed, spread, d);
        const cudaError_t e = cudaMemcpy(
            mismatches, d, 2 * sizeof *d, cudaMemcpyDeviceToHost);
        cudaFree(d);
        return (e == cudaSuccess) ? PUMAS_RETURN_SUCCESS :
                                    PUMAS_RETURN_INTERNAL_ERROR;
}

/* Measured FP64 FMA peak of the current device in TFLOP/s (2 flop per FMA). */
enum pumas_return pumas_b200_fp64_peak(double * tflops)
{
        if (tflops == nullptr) return PUMAS_RETURN_VALUE_ERROR;
        *tflops = 0.;
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) != cudaSuccess)
                return PUMAS_RETURN_INTERNAL_ERROR;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        double * d_out = nullptr;
        if (cudaMalloc(&d_out, sizeof(double)) != cudaSuccess)
                return PUMAS_RETURN_INTERNAL_ERROR;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        const int threads = 512, blocks = sms * 4, iters = 1 << 16;
        double best = 0.;
        for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0);
                fp64_peak_kernel<<<blocks, threads>>>(
                    d_out, iters, 1.0000001, 1E-09);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms = 0.f;
                cudaEventElapsedTime(&ms, e0, e1);
                const double flops =
                    2. * 8. * (double)iters * (double)threads * (double)blocks;
                const double t = flops / ((double)ms * 1E-03) * 1E-12;
                if ((rep > 0) && (t > best)) best = t;
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        cudaFree(d_out);
        if (cudaGetLastError() != cudaSuccess)
                return PUMAS_RETURN_INTERNAL_ERROR;
        *tflops = best;
        return PUMAS_RETURN_SUCCESS;
}

} /* extern "C" */
