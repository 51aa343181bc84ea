// fp64_issue.cu -- what the FP64 pipe of one sm_100a scheduler can take, measured: the ceilings DESIGN.md and profiles/README.md
// quote for the FP64-bound kernels come from here.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue fp64_issue.cu
//   latency    : one dependent DFMA chain, 1 warp per scheduler
//   dfma       : 8 independent DFMA chains per thread
//   dfma+imad  : the same plus 8 independent IMAD chains (does an FP64 instruction hold the issue port for two cycles?)
//   dfma+2imad : two integer instructions per DFMA
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double *out, int iters, double a, double b, int ia)
{
    double x[8]; int y[16];
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 16; i++) y[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int u = 0; u < 8; u++) x[0] = fma(x[0], a, b);
        } else {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                x[u] = fma(x[u], a, b);
                if (MODE >= 2) y[u] = y[u] * ia + 7;
                if (MODE >= 3) y[8 + u] = y[8 + u] * ia + 3;
            }
        }
    }
    long long t1 = clock64();
    double s = 0; int si = 0;
    for (int i = 0; i < 8; i++) s += x[i];
    for (int i = 0; i < 16; i++) si += y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + si;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (double)(t1 - t0);
}

template <int MODE>
void run(const char *name, int threads, double *d)
{
    const int iters = 4096;
    k<MODE><<<1, threads>>>(d, iters, 1.0000001, 1e-9, 3);
    cudaDeviceSynchronize();
    k<MODE><<<1, threads>>>(d, iters, 1.0000001, 1e-9, 3);
    double cyc;
    cudaMemcpy(&cyc, d + threads, sizeof(double), cudaMemcpyDeviceToHost);
    const int warps_per_sched = threads / 128;
    const double dfma = 8.0 * iters * warps_per_sched;
    printf("%-12s warps/scheduler %d: %.0f cycles, %.3f cycles per DFMA warp-instruction per scheduler\n", name, warps_per_sched, cyc, cyc / dfma);
}

int main()
{
    double *d; cudaMalloc(&d, sizeof(double) * 2048);
    for (int threads : {128, 256, 512, 1024}) {
        run<0>("latency", threads, d);
        run<1>("dfma", threads, d);
        run<2>("dfma+imad", threads, d);
        run<3>("dfma+2imad", threads, d);
    }
    return cudaGetLastError() != cudaSuccess;
}
