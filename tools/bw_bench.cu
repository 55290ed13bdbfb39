// Micro-benchmark: steady-state read bandwidth from HBM and from L2, plain LDG vs a TMA ring.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int kThreads = 256;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, unsigned n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, unsigned par) {
    unsigned ok; const uint32_t a = smem_u32(b);
    do { asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(par) : "memory"); } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__global__ void k_ldg(const double2 *a, size_t n, int reps, double *out)
{
    double s = 0;
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { const double2 v = a[i]; s += v.x + v.y; }
    if (s == 1.2345) out[0] = s;
}
// ring of STAGES tiles of TILE_BYTES each, one CTA per SM
template <int STAGES, int TILE_BYTES>
__global__ void __launch_bounds__(kThreads, 1) k_tma(const char *a, size_t bytes, int reps, double *out)
{
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ __align__(8) uint64_t full[STAGES];
    const int tid = threadIdx.x, G = gridDim.x;
    const long long NT = bytes / TILE_BYTES;
    const long long n_my = (blockIdx.x < NT) ? (NT - 1 - blockIdx.x) / G + 1 : 0;
    const long long total = n_my * reps;
    if (tid == 0) { for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    auto issue = [&](long long k) {
        const int s = (int)(k % STAGES);
        const long long tile = blockIdx.x + (k % n_my) * G;
        mbar_expect_tx(&full[s], TILE_BYTES);
        bulk_g2s(raw + (size_t)s * TILE_BYTES, a + tile * TILE_BYTES, TILE_BYTES, &full[s]);
    };
    if (tid == 0) for (long long k = 0; k < (total < STAGES ? total : STAGES); ++k) issue(k);
    double s = 0;
    for (long long k = 0; k < total; ++k) {
        const int sg = (int)(k % STAGES);
        mbar_wait(&full[sg], (unsigned)((k / STAGES) & 1));
        const double2 *p = reinterpret_cast<const double2 *>(raw + (size_t)sg * TILE_BYTES);
        for (int j = tid; j < TILE_BYTES / 16; j += kThreads) { const double2 v = p[j]; s += v.x + v.y; }
        __syncthreads();
        if (tid == 0 && k + STAGES < total) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(k + STAGES); }
    }
    if (s == 1.2345) out[0] = s;
}
template <int STAGES, int TILE_BYTES>
float run_tma(const char *a, size_t bytes, int reps, double *out)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t smem = (size_t)STAGES * TILE_BYTES;
    cudaFuncSetAttribute(k_tma<STAGES, TILE_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_tma<STAGES, TILE_BYTES><<<148, kThreads, smem>>>(a, bytes, 1, out);
    cudaEventRecord(e0);
    k_tma<STAGES, TILE_BYTES><<<148, kThreads, smem>>>(a, bytes, reps, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}
int main()
{
    char *a; double *out;
    const size_t big = 1024ull << 20;
    cudaMalloc(&a, big); cudaMalloc(&out, 8); cudaMemset(a, 0, big);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    struct { const char *name; size_t bytes; int reps; } cfg[] = {{"HBM 1024 MB x1", big, 1}, {"L2 32 MB x32", 32ull << 20, 32}, {"L2 64 MB x16", 64ull << 20, 16}, {"L2 96 MB x10", 96ull << 20, 10}, {"116 MB x8", 116ull << 20, 8}, {"133 MB x8", 133ull << 20, 8}};
    for (auto &c : cfg) {
        k_ldg<<<148 * 8, 256>>>((const double2 *)a, c.bytes / 16, 1, out);
        cudaEventRecord(e0); k_ldg<<<148 * 8, 256>>>((const double2 *)a, c.bytes / 16, c.reps, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("LDG 1184x256   %-16s %8.1f us  %6.0f GB/s\n", c.name, ms * 1e3, (double)c.bytes * c.reps / (ms * 1e-3) / 1e9);
        ms = run_tma<8, 14336>(a, c.bytes, c.reps, out);
        printf("TMA 8x14KB     %-16s %8.1f us  %6.0f GB/s\n", c.name, ms * 1e3, (double)(c.bytes / 14336 * 14336) * c.reps / (ms * 1e-3) / 1e9);
        ms = run_tma<4, 32768>(a, c.bytes, c.reps, out);
        printf("TMA 4x32KB     %-16s %8.1f us  %6.0f GB/s\n", c.name, ms * 1e3, (double)(c.bytes / 32768 * 32768) * c.reps / (ms * 1e-3) / 1e9);
        ms = run_tma<12, 16384>(a, c.bytes, c.reps, out);
        printf("TMA 12x16KB    %-16s %8.1f us  %6.0f GB/s\n", c.name, ms * 1e3, (double)(c.bytes / 16384 * 16384) * c.reps / (ms * 1e-3) / 1e9);
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
