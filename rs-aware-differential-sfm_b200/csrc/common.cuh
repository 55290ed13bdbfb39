// common.cuh -- context, error handling, scratch management and block reductions shared by the
// sm_100a kernels behind include/rsdsfm.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/rsdsfm.h"

namespace rsdsfm {

constexpr int kNumSMsB200 = 148;
#ifndef RS_THREADS
#define RS_THREADS 256
#endif
constexpr int kThreads = RS_THREADS;   // threads per CTA for the per-pixel map-reduce kernels (RS_THREADS: experiments only)
constexpr int kWarps = kThreads / 32;

// A grow-only device buffer owned by the context.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

}  // namespace rsdsfm

constexpr int kMaxLanes = 16;

struct rsdsfm_ctx {
    int device = 0;
    int num_sms = rsdsfm::kNumSMsB200;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    long long launches = 0;
    std::string err;
    // scratch
    std::vector<rsdsfm::DevBuf *> bufs;
    rsdsfm::DevBuf partials, sums, pix, dA, dB, rdepth, misc, stage[16], winner, poses;
    rsdsfm::DevBuf hyp, rpart, scan, lm_shared, exc, splat_tab, flow_t;
    rsdsfm::DevBuf pipe[16];  // intermediates of the fused a2-a15 driver (pipeline.cu)
    int exc_cap = 0;          // capacity (entries) of the clamped-pixel exception list
    // Row split of one solve over several GPUs (refine.cu / lm_kernel.cuh): this GPU's mailbox (its own allocation, so
    // that it can be exported through CUDA IPC) and the mailboxes of the group, peer_mail[my_peer] == mailbox.
    void *mailbox = nullptr;
    void *peer_mail[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool peer_ipc[8] = {false, false, false, false, false, false, false, false};   // opened with cudaIpcOpenMemHandle
    int n_peers = 1, my_peer = 0;
    unsigned int peer_epoch = 0;   // split solves since the group was formed (identical on every GPU of the group)
    void *pinned = nullptr;   // small pinned host buffer for reduced sums / scalars
    size_t pinned_cap = 0;
    // Pipelined sequences (pipeline.cu): while pair i computes on `stream`, pair i+1 uploads on
    // `s_in` and pair i-1 downloads on `s_out`.  Two I/O slots: staging buffers stage[8*slot+j]
    // and a pinned slot area (LM control block in/out + depth statistics) each.
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_cdone[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    int io_slot = 0;
    // Multi-lane sequences: the LM solve leaves every SM idle for a quarter of each iteration (grid exchange +
    // serial controller), and that part does not shrink with the grid.  rsdsfm_refine_rectify_sequence therefore
    // runs several pairs at once, each on a `lane` (lane 0 = this context, lanes[k-1] = full extra contexts with
    // their own streams and buffers), each solve on a FRACTION of the SMs (lm_grid CTAs): per LM iteration
    // 53.8 us on 148 SMs, 86.0 on 74, 156.6 on 37 -- i.e. 53.8 / 43.0 / 39.1 us of whole-GPU time.  More lanes
    // exist than solves fit on the GPU: a lane that is uploading, downloading or waiting for the host costs no SM.
    rsdsfm_ctx *lanes[kMaxLanes - 1] = {};
    int lm_grid = 0;             // CTAs of the persistent LM kernel; 0 = one per SM
    void *pinned_io = nullptr;   // 2 x kPinnedSlotBytes, allocated with the context
    // per-kernel profiling (bench.py's roofline): CUDA events around the LM passes
    bool profile = false;
    cudaEvent_t pe0[2] = {nullptr, nullptr}, pe1[2] = {nullptr, nullptr};   // per I/O slot
    double prof_detail[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // ms: pass A pixel loop / CTA reduce / controller, same for pass B
    double prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // [0] pass A ms, [1] pass A phases, [2] pass A residual blocks,
                                                 // [3] pass B ms, [4] pass B phases, [5] pass B residual blocks,
                                                 // [6] LM kernel ms (CUDA events), [7] LM kernel launches
};

namespace rsdsfm {

extern thread_local std::string g_create_error;

// pinned slot area: [0, 8192) control block read back, [8192, 16128) initial control block,
// [16128, 16384) depth statistics
constexpr size_t kPinnedSlotBytes = 16384;
inline char *pinned_slot(rsdsfm_ctx *ctx) { return (char *)ctx->pinned_io + kPinnedSlotBytes * (size_t)ctx->io_slot; }
inline void *pinned_lm_result(rsdsfm_ctx *ctx) { return pinned_slot(ctx); }
inline void *pinned_lm_init(rsdsfm_ctx *ctx) { return pinned_slot(ctx) + 8192; }
inline double *pinned_stats(rsdsfm_ctx *ctx) { return (double *)(pinned_slot(ctx) + kPinnedSlotBytes - 256); }

inline int fail(rsdsfm_ctx *ctx, int code, const char *what, cudaError_t e = cudaSuccess)
{
    char buf[512];
    if (e != cudaSuccess)
        snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    else
        snprintf(buf, sizeof buf, "%s", what);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

#define RS_CUDA(ctx, call)                                                        \
    do {                                                                          \
        cudaError_t e__ = (call);                                                 \
        if (e__ != cudaSuccess) return ::rsdsfm::fail((ctx), RSDSFM_ERR_CUDA, #call, e__); \
    } while (0)

#define RS_TRY(expr)                         \
    do {                                     \
        int rc__ = (expr);                   \
        if (rc__ != RSDSFM_OK) return rc__;  \
    } while (0)

inline int ensure(rsdsfm_ctx *ctx, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap && b.p) return RSDSFM_OK;
    RS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (b.p) { RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t cap = bytes < 256 ? 256 : bytes;
    cudaError_t e = cudaMalloc(&b.p, cap);
    if (e != cudaSuccess) return fail(ctx, RSDSFM_ERR_NOMEM, "cudaMalloc", e);
    b.cap = cap;
    return RSDSFM_OK;
}

inline int ensure_pinned(rsdsfm_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->pinned_cap) return RSDSFM_OK;
    if (ctx->pinned) { cudaStreamSynchronize(ctx->stream); cudaFreeHost(ctx->pinned); ctx->pinned = nullptr; }
    RS_CUDA(ctx, cudaMallocHost(&ctx->pinned, bytes));
    ctx->pinned_cap = bytes;
    return RSDSFM_OK;
}

// Stage an input array: returns a device pointer holding `bytes` of `src` (which lives in `mem`).
inline int stage_in(rsdsfm_ctx *ctx, int mem, int slot, const void *src, size_t bytes, const void **dev,
                    cudaStream_t on = nullptr)
{
    if (mem == RSDSFM_DEVICE || src == nullptr) { *dev = src; return RSDSFM_OK; }
    RS_TRY(ensure(ctx, ctx->stage[slot], bytes));
    RS_CUDA(ctx, cudaMemcpyAsync(ctx->stage[slot].p, src, bytes, cudaMemcpyHostToDevice, on ? on : ctx->stream));
    *dev = ctx->stage[slot].p;
    return RSDSFM_OK;
}
// Reserve an output array: a device pointer of `bytes`; call stage_out afterwards.
inline int stage_out_reserve(rsdsfm_ctx *ctx, int mem, int slot, void *dst, size_t bytes, void **dev)
{
    if (mem == RSDSFM_DEVICE || dst == nullptr) { *dev = dst; return RSDSFM_OK; }
    RS_TRY(ensure(ctx, ctx->stage[slot], bytes));
    *dev = ctx->stage[slot].p;
    return RSDSFM_OK;
}
inline int stage_out(rsdsfm_ctx *ctx, int mem, void *dst, const void *dev, size_t bytes, cudaStream_t on = nullptr)
{
    if (mem == RSDSFM_DEVICE || dst == nullptr || bytes == 0) return RSDSFM_OK;
    RS_CUDA(ctx, cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, on ? on : ctx->stream));
    return RSDSFM_OK;
}

inline int grid_for(const rsdsfm_ctx *ctx, long long n, int per_sm = 2)
{
    long long need = (n + kThreads - 1) / kThreads;
    long long cap = (long long)ctx->num_sms * per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

#ifdef __CUDACC__
// ---- block reductions -------------------------------------------------------------------------
// NS sums + NM maxima per thread -> one row of `out` per CTA: out[blockIdx.x * (NS+NM) + j].
// Fixed shuffle tree + fixed warp order: bit-reproducible for a given launch geometry.
template <int NS, int NM>
__device__ __forceinline__ void block_reduce_store(double (&s)[NS > 0 ? NS : 1], double (&mx)[NM > 0 ? NM : 1],
                                                   double *out, int row = -1)
{
    if (row < 0) row = (int)blockIdx.x;
    __shared__ double sh[kWarps][NS + NM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        double v = s[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sh[warp][j] = v;
    }
#pragma unroll
    for (int j = 0; j < NM; ++j) {
        double v = mx[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) sh[warp][NS + j] = v;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < NS + NM; j += blockDim.x) {
        double v = sh[0][j];
        if (j < NS) { for (int w2 = 1; w2 < kWarps; ++w2) v += sh[w2][j]; }
        else        { for (int w2 = 1; w2 < kWarps; ++w2) v = fmax(v, sh[w2][j]); }
        out[(size_t)row * (NS + NM) + j] = v;
    }
}

// NaN-propagating "bad value" detector usable inside fmax reductions (fmax drops NaNs).
__device__ __forceinline__ double bad_flag(double a) { return isfinite(a) ? 0.0 : 1.0; }
#endif

// out[j] = reduce over blocks b of partials[b*(ns+nm)+j]; sums in ascending block order.
void launch_final_reduce(rsdsfm_ctx *ctx, const double *partials, int nblocks, int ns, int nm, double *out);

}  // namespace rsdsfm
