// Micro-benchmark: FP64 DFMA dependent-issue latency and pipe throughput on the current GPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_chain(double *out, long long *cycles, int iters, double a, double b)
{
    double x[ILP];
    for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x * 1e-3 + j;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < ILP; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int ILP>
void run(int warps_per_sm, int iters)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(double) * 148 * 1024);
    cudaMalloc(&cyc, 8);
    k_chain<ILP><<<148, warps_per_sm * 32>>>(out, cyc, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    k_chain<ILP><<<148, warps_per_sm * 32>>>(out, cyc, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per = (double)h / iters;            // cycles per loop iteration (ILP DFMAs per warp)
    double per_smsp_instr = per / (ILP * (warps_per_sm / 4.0));   // cycles per warp-DFMA per SMSP
    printf("warps/SM=%2d ILP=%d: %.2f cycles/iter, %.2f cycles per warp-DFMA per SMSP\n", warps_per_sm, ILP, per, warps_per_sm >= 4 ? per_smsp_instr : per / ILP);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    const int it = 20000;
    run<1>(4, it); run<2>(4, it); run<4>(4, it); run<8>(4, it); run<16>(4, it);
    run<1>(8, it); run<2>(8, it); run<4>(8, it); run<8>(8, it);
    run<1>(16, it); run<4>(16, it);
    run<1>(32, it); run<2>(32, it);
    return 0;
}
