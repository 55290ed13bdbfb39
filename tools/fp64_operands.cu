// Micro-benchmark: FP64 issue rate as a function of how many DISTINCT register operands an instruction reads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_operands fp64_operands.cu
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: acc = fma(acc, a, b)            one varying operand (a, b stay in the operand reuse cache)
// MODE 1: acc_i = fma(u_i, v_i, acc_i)    three distinct register pairs per instruction
// MODE 2: rank-1 update of a packed 7x7 triangle + 7-vector (the solver's pattern: 35 DFMAs, k_j reused along a row)
// MODE 3: acc_i = fma(u_i, u_i, acc_i)    two distinct
// MODE 4: acc_i = u_i * v_i (DMUL, two distinct, result chained through an add every 8th)
template <int MODE>
__global__ void k(double *out, long long *cycles, int iters, const double *in)
{
    double u[8], v[8], acc[36];
    for (int j = 0; j < 8; ++j) { u[j] = in[j] + threadIdx.x * 1e-9; v[j] = in[8 + j] - threadIdx.x * 1e-9; }
    for (int j = 0; j < 36; ++j) acc[j] = 0.0;
    const double a = in[16], b = in[17];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = fma(acc[j], a, b);
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = fma(u[j & 7], v[(j + (j >> 3)) & 7], acc[j]);
        } else if (MODE == 2) {
            int t = 0;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                acc[28 + j] = fma(u[j], u[7], acc[28 + j]);
#pragma unroll
                for (int c = j; c < 7; ++c, ++t) acc[t] = fma(u[j], u[c], acc[t]);
            }
        } else if (MODE == 3) {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = fma(u[j & 7], u[j & 7], acc[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = u[j & 7] * v[(j + (j >> 3)) & 7] + acc[j] * 0.0;
        }

    }
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < 36; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(const char *name, int warps_per_sm, int iters, int ninstr, const double *in)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(double) * 148 * 1024);
    cudaMalloc(&cyc, 8);
    for (int r = 0; r < 2; ++r) { k<MODE><<<148, warps_per_sm * 32>>>(out, cyc, iters, in); cudaDeviceSynchronize(); }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s warps/SM=%2d: %.2f cycles per warp-instruction per SMSP\n", name, warps_per_sm, (double)h / iters / (ninstr * (warps_per_sm / 4.0)));
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    double h[18]; for (int j = 0; j < 18; ++j) h[j] = 1.0 + 0.01 * j;
    double *in; cudaMalloc(&in, sizeof h); cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    const int it = 4000;
    for (int w = 4; w <= 8; w += 4) {
        run<0>("fma(acc, a, b)  [1 varying operand]", w, it, 32, in);
        run<1>("fma(u_i, v_k, acc_j)  [3 distinct]", w, it, 32, in);
        run<3>("fma(u_i, u_i, acc_j)  [2 distinct]", w, it, 32, in);
        run<2>("rank-1 triangle + vector (35 DFMA)", w, it, 35, in);
    }
    return 0;
}
