// Micro-benchmark: does integer / move work issued between FP64 instructions cost FP64 throughput?
// Each loop trip: 32 DFMAs (2 or 3 distinct register operands) interleaved with NI integer LOP3/IADD per DFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NI, int DISTINCT>
__global__ void k(double *out, long long *cycles, int iters, const double *in, const int *iin)
{
    double u[8], v[8], acc[32];
    int z[8];
    for (int j = 0; j < 8; ++j) { u[j] = in[j] + threadIdx.x * 1e-9; v[j] = in[8 + j] - threadIdx.x * 1e-9; z[j] = iin[j] + threadIdx.x; }
    for (int j = 0; j < 32; ++j) acc[j] = 0.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            acc[j] = (DISTINCT == 3) ? fma(u[j & 7], v[(j + (j >> 3)) & 7], acc[j]) : fma(u[j & 7], u[j & 7], acc[j]);
#pragma unroll
            for (int q = 0; q < NI; ++q) z[(j + q) & 7] = (z[(j + q) & 7] ^ z[(j + q + 3) & 7]) + z[(j + q + 5) & 7];
        }
    }
    long long t1 = clock64();
    double s = 0; int zi = 0;
    for (int j = 0; j < 32; ++j) s += acc[j];
    for (int j = 0; j < 8; ++j) zi += z[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + zi;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int NI, int DISTINCT>
void run(int warps_per_sm, const double *in, const int *iin)
{
    const int iters = 4000;
    double *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(double) * 148 * 1024); cudaMalloc(&cyc, 8);
    for (int r = 0; r < 2; ++r) { k<NI, DISTINCT><<<148, warps_per_sm * 32>>>(out, cyc, iters, in, iin); cudaDeviceSynchronize(); }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA with %d distinct register operands + %d integer instructions each, %d warps/SM: %.2f cycles per DFMA per SMSP\n",
           DISTINCT, 2 * NI, warps_per_sm, (double)h / iters / (32 * (warps_per_sm / 4.0)));
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    double h[16]; int hi[8];
    for (int j = 0; j < 16; ++j) h[j] = 1.0 + 0.01 * j;
    for (int j = 0; j < 8; ++j) hi[j] = 3 * j + 1;
    double *in; int *iin;
    cudaMalloc(&in, sizeof h); cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    cudaMalloc(&iin, sizeof hi); cudaMemcpy(iin, hi, sizeof hi, cudaMemcpyHostToDevice);
    for (int w = 4; w <= 8; w += 4) {
        if (w == 4) { run<0, 2>(4, in, iin); run<1, 2>(4, in, iin); run<2, 2>(4, in, iin); run<0, 3>(4, in, iin); run<1, 3>(4, in, iin); run<2, 3>(4, in, iin); }
        else        { run<0, 2>(8, in, iin); run<1, 2>(8, in, iin); run<2, 2>(8, in, iin); run<0, 3>(8, in, iin); run<1, 3>(8, in, iin); run<2, 3>(8, in, iin); }
    }
    return 0;
}
