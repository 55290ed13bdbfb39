// Micro-benchmark: how fast can 148 persistent CTAs stream the LM working set (3 double2 arrays +
// 1 double array, 2,073,600 entries) from HBM?  (a) plain LDG.128 grid-stride, many CTAs;
// (b) the solver's pattern: one CTA per SM, TMA bulk copies into an 8-stage shared-memory ring.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kThreads = 256, kTile = 256, kStages = 8;
struct Stage { double2 xy[kTile], uu[kTile], aa[kTile]; double d[kTile]; };

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

__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, unsigned bytes, uint64_t *bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

// same ring, but tiles below `pin_tiles` are fetched with an L2 evict_last policy, the rest evict_first
__global__ void __launch_bounds__(kThreads, 1) k_tma_pin(const double2 *xy, const double2 *uu, const double2 *aa, const double *d, int m, double *out, int pin_tiles)
{
    extern __shared__ __align__(128) unsigned char raw[];
    Stage *st = reinterpret_cast<Stage *>(raw);
    __shared__ __align__(8) uint64_t full[kStages];
    const int tid = threadIdx.x, G = gridDim.x;
    const int NT = (m + kTile - 1) / kTile;
    const int n_my = ((int)blockIdx.x < NT) ? (NT - 1 - (int)blockIdx.x) / G + 1 : 0;
    uint64_t pol_last, pol_first;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    if (tid == 0) { for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    auto issue = [&](int k) {
        const int s = k % kStages, tile = blockIdx.x + k * G, first = tile * kTile;
        const int cnt = min(kTile, m - first);
        const uint64_t pol = tile < pin_tiles ? pol_last : pol_first;
        mbar_expect_tx(&full[s], cnt * 56);
        bulk_g2s_hint(st[s].xy, xy + first, cnt * 16, &full[s], pol); bulk_g2s_hint(st[s].uu, uu + first, cnt * 16, &full[s], pol);
        bulk_g2s_hint(st[s].aa, aa + first, cnt * 16, &full[s], pol); bulk_g2s_hint(st[s].d, d + first, cnt * 8, &full[s], pol_first);
    };
    if (tid == 0) for (int k = 0; k < min(kStages, n_my); ++k) issue(k);
    double s = 0;
    for (int k = 0; k < n_my; ++k) {
        const int sg = k % kStages;
        mbar_wait(&full[sg], (k / kStages) & 1);
        const int i = (blockIdx.x + k * G) * kTile + tid;
        double v = 0;
        if (i < m) { const double2 a = st[sg].xy[tid], b = st[sg].uu[tid], c = st[sg].aa[tid]; v = a.x + a.y + b.x + b.y + c.x + c.y + st[sg].d[tid]; }
        __syncthreads();
        if (tid == 0 && k + kStages < n_my) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(k + kStages); }
        s += v;
    }
    if (s == 1.2345) out[0] = s;
}

__global__ void k_ldg(const double2 *xy, const double2 *uu, const double2 *aa, const double *d, int m, double *out)
{
    double s = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const double2 a = xy[i], b = uu[i], c = aa[i];
        s += a.x + a.y + b.x + b.y + c.x + c.y + d[i];
    }
    if (s == 1.2345) out[0] = s;
}

__global__ void __launch_bounds__(kThreads, 1) k_tma(const double2 *xy, const double2 *uu, const double2 *aa, const double *d, int m, double *out, int write_back, double *dw)
{
    extern __shared__ __align__(128) unsigned char raw[];
    Stage *st = reinterpret_cast<Stage *>(raw);
    __shared__ __align__(8) uint64_t full[kStages];
    const int tid = threadIdx.x, G = gridDim.x;
    const int NT = (m + kTile - 1) / kTile;
    const int n_my = ((int)blockIdx.x < NT) ? (NT - 1 - (int)blockIdx.x) / G + 1 : 0;
    if (tid == 0) { for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    auto issue = [&](int k) {
        const int s = k % kStages, first = (blockIdx.x + k * G) * kTile;
        const int cnt = min(kTile, m - first);
        mbar_expect_tx(&full[s], cnt * 56);
        bulk_g2s(st[s].xy, xy + first, cnt * 16, &full[s]); bulk_g2s(st[s].uu, uu + first, cnt * 16, &full[s]);
        bulk_g2s(st[s].aa, aa + first, cnt * 16, &full[s]); bulk_g2s(st[s].d, d + first, cnt * 8, &full[s]);
    };
    if (tid == 0) for (int k = 0; k < min(kStages, n_my); ++k) issue(k);
    double s = 0;
    for (int k = 0; k < n_my; ++k) {
        const int sg = k % kStages;
        mbar_wait(&full[sg], (k / kStages) & 1);
        const int i = (blockIdx.x + k * G) * kTile + tid;
        double v = 0;
        if (i < m) { const double2 a = st[sg].xy[tid], b = st[sg].uu[tid], c = st[sg].aa[tid]; v = a.x + a.y + b.x + b.y + c.x + c.y + st[sg].d[tid]; }
        __syncthreads();
        if (tid == 0 && k + kStages < n_my) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(k + kStages); }
        s += v;
        if (write_back && i < m) dw[i] = v;
    }
    if (s == 1.2345) out[0] = s;
}

int main()
{
    const int m = 2073600;
    double2 *xy, *uu, *aa; double *d, *dw, *out;
    cudaMalloc(&xy, 16ull * m); cudaMalloc(&uu, 16ull * m); cudaMalloc(&aa, 16ull * m); cudaMalloc(&d, 8ull * m + 16); cudaMalloc(&dw, 8ull * m); cudaMalloc(&out, 8);
    cudaMemset(xy, 0, 16ull * m); cudaMemset(uu, 0, 16ull * m); cudaMemset(aa, 0, 16ull * m); cudaMemset(d, 0, 8ull * m);
    // a 256 MB buffer written between runs evicts the working set from L2
    char *flush; cudaMalloc(&flush, 256u << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t smem = sizeof(Stage) * kStages;
    cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const double bytes = 56.0 * m;
    for (int variant = 0; variant < 5; ++variant) {
        float best = 1e9f;
        for (int rep = 0; rep < 6; ++rep) {
            cudaMemset(flush, rep, 256u << 20);
            cudaEventRecord(e0);
            if (variant == 0) k_ldg<<<148 * 8, 256>>>(xy, uu, aa, d, m, out);
            if (variant == 1) k_ldg<<<148 * 2, 1024>>>(xy, uu, aa, d, m, out);
            if (variant == 2) k_ldg<<<148, 256>>>(xy, uu, aa, d, m, out);
            if (variant == 3) k_tma<<<148, kThreads, smem>>>(xy, uu, aa, d, m, out, 0, dw);
            if (variant == 4) k_tma<<<148, kThreads, smem>>>(xy, uu, aa, d, m, out, 1, dw);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        const char *names[] = {"LDG 1184 CTAs x256", "LDG 296 CTAs x1024", "LDG 148 CTAs x256 (8 warps/SM)", "TMA ring 148 CTAs (read only)", "TMA ring + 8 B/px write"};
        printf("%-34s %.1f us  %.0f GB/s\n", names[variant], best * 1e3, (bytes + (variant == 4 ? 8.0 * m : 0)) / (best * 1e-3) / 1e9);
    }
    cudaFuncSetAttribute(k_tma_pin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int NT = (m + kTile - 1) / kTile;
    for (int pct = 0; pct <= 100; pct += 10) {
        cudaMemset(flush, 1, 256u << 20);
        float tot = 0; int cnt = 0;
        for (int rep = 0; rep < 12; ++rep) {
            cudaEventRecord(e0);
            k_tma_pin<<<148, kThreads, smem>>>(xy, uu, aa, d, m, out, NT * pct / 100);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep >= 4) { tot += ms; cnt++; }
        }
        printf("pinned %3d%% of xy/uu/aa (%5.1f MB): %.1f us per pass (back-to-back, no flush)  eff %.0f GB/s\n", pct, 48.0 * m * pct / 100 / 1e6, tot / cnt * 1e3, bytes / (tot / cnt * 1e-3) / 1e9);
    }
    // no hints at all, back-to-back
    { float tot = 0; int cnt = 0;
      for (int rep = 0; rep < 12; ++rep) { cudaEventRecord(e0); k_tma<<<148, kThreads, smem>>>(xy, uu, aa, d, m, out, 0, dw); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep >= 4) { tot += ms; cnt++; } }
      printf("no hints, back-to-back: %.1f us per pass\n", tot / cnt * 1e3); }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
