// dmma_bench.cu -- does the FP64 tensor-core path (mma.sync m8n8k4 f64) run beside the FP64 FMA
// pipe on B200, and at what rate?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NF, int NM>   // per loop iteration: NF independent DFMAs and NM independent DMMAs per warp
__global__ void __launch_bounds__(256) k_mix(double *out, int iters, long long *cycles)
{
    double f[NF > 0 ? NF : 1], c[NM > 0 ? 2 * NM : 1];
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * (threadIdx.x + 1);
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) f[i] = i;
    for (int i = 0; i < (NM > 0 ? 2 * NM : 1); ++i) c[i] = i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
#pragma unroll
        for (int i = 0; i < NM; ++i) dmma(c[2 * i], c[2 * i + 1], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) s += f[i];
    for (int i = 0; i < (NM > 0 ? 2 * NM : 1); ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NF, int NM>
void run(const char *name, double *out, long long *cyc, int threads)
{
    const int iters = 20000;
    k_mix<NF, NM><<<148, threads>>>(out, 100, cyc);
    cudaDeviceSynchronize();
    k_mix<NF, NM><<<148, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < 148; ++i) mean += (double)h[i];
    mean /= 148;
    const int warps_per_smsp = threads / 32 / 4;
    const double per_iter = mean / iters;   // cycles per loop iteration (all warps of an SMSP run concurrently)
    printf("%-28s threads %3d  cycles/iter %8.2f  -> per SMSP: %.2f cyc per warp-DFMA, %.2f cyc per warp-DMMA (if alone)  flops/clk/SM %.1f\n",
           name, threads, per_iter, NF ? per_iter / (NF * warps_per_smsp) : 0.0, NM ? per_iter / (NM * warps_per_smsp) : 0.0,
           (NF * 64.0 + NM * 512.0) * (threads / 32) / per_iter);
}

int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, sizeof(double) * 148 * 1024);
    cudaMalloc(&cyc, sizeof(long long) * 148);
    for (int threads : {128, 256, 512}) {
        if (threads == 128) { run<16, 0>("16 DFMA", out, cyc, 128); run<0, 8>("8 DMMA", out, cyc, 128); run<16, 4>("16 DFMA + 4 DMMA", out, cyc, 128); run<16, 8>("16 DFMA + 8 DMMA", out, cyc, 128); }
        if (threads == 256) { run<16, 0>("16 DFMA", out, cyc, 256); run<0, 8>("8 DMMA", out, cyc, 256); run<16, 4>("16 DFMA + 4 DMMA", out, cyc, 256); run<16, 8>("16 DFMA + 8 DMMA", out, cyc, 256); run<16, 2>("16 DFMA + 2 DMMA", out, cyc, 256); }
        if (threads == 512) { run<16, 0>("16 DFMA", out, cyc, 512); run<0, 8>("8 DMMA", out, cyc, 512); run<16, 4>("16 DFMA + 4 DMMA", out, cyc, 512); }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
