// FP64 micro-benchmarks: dependent-chain latency and per-SM throughput vs warps.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void chain(double *out, long long *cyc, double a, double b, int iters)
{
    double x = a + threadIdx.x * 1e-9, y = b;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (OP == 0) x = __dadd_rn(x, y);
            if (OP == 1) x = __dmul_rn(x, y);
            if (OP == 2) x = __fma_rn(x, y, y);
            if (OP == 3) x = 1.0 / x + y;
            if (OP == 4) x = sqrt(x) + y;
            if (OP == 5) x = log(x) + y;
            if (OP == 6) { double s, c; sincos(x, &s, &c); x = s + c; }
            if (OP == 7) x = __int_as_float(__float_as_int((float)x) + 1) ;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// throughput: ILP independent chains per thread
template <int ILP> __global__ void tput(double *out, double a, double b, int iters)
{
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = a + k + threadIdx.x * 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int k = 0; k < ILP; k++) x[k] = __fma_rn(x[k], b, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; k++) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 8 * 4); cudaMalloc(&cyc, 8);
    const char *names[] = {"dadd", "dmul", "dfma", "div+add", "sqrt+add", "log+add", "sincos", "cvt"};
    long long h;
    int iters = 1000;
#define RUN(OP) chain<OP><<<1, 32>>>(out, cyc, 1.5, 1.0000001, iters); chain<OP><<<1, 32>>>(out, cyc, 1.5, 1.0000001, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-10s dependent: %.1f cycles/op\n", names[OP], (double)h / (iters * 16.0));
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 1; warps <= 32; warps *= 2) {
        float ms;
        int it = 20000;
#define TP(ILP) tput<ILP><<<148, warps * 32>>>(out, 1.5, 0.999999, 100); cudaEventRecord(e0); tput<ILP><<<148, warps * 32>>>(out, 1.5, 0.999999, it); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); printf("warps/SM %2d ILP %d: %.2f TFLOP/s (DFMA=2)\n", warps, ILP, 2.0 * 148 * warps * 32 * (double)it * 8 * ILP / (ms * 1e-3) / 1e12);
        TP(1) TP(2) TP(4)
    }
    printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
