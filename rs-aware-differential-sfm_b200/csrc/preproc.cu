// preproc.cu -- flow flattening / normalisation, RS scale factors and consensus-set gathering.
// Compiled with --fmad=false (results are bit-identical to the reference's scalar arithmetic).
//
//   a2  main.cc:398-432, errorMeasure.cpp:66-97      k_flat_*   (stream compaction in COLUMN-MAJOR
//                                                     pixel order: count -> scan -> scatter)
//   a3  minimal::getAlpha    minimal.cc:179-186       k_alpha
//   a4  minimal::getAlphaK   minimal.cc:188-197       k_alpha
//   a6  tail of minimal::ransac  minimal.cc:291-305   k_gather_*  (ascending point index)
#include "common.cuh"
#include "lm_layout.h"

namespace rsdsfm {

constexpr int kChunk = kThreads * 4;   // points per CTA in the compaction kernels

// Exclusive rank of this thread's flag among the flags of the CTA, in thread order.
// Returns the CTA total through `total`.
__device__ __forceinline__ int block_rank(bool flag, int &total)
{
    __shared__ int warp_cnt[kWarps + 1];
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int within = __popc(bal & ((1u << lane) - 1u));
    __syncthreads();
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int base = 0, tot = 0;
    for (int w2 = 0; w2 < kWarps; ++w2) { const int c = warp_cnt[w2]; if (w2 < warp) base += c; tot += c; }
    total = tot;
    return base + within;
}

__device__ __forceinline__ bool flow_kept(const double *__restrict__ flow_img, int rows, int cols, long long p,
                                          long long total, double thr, double &dx, double &dy, int &i, int &j)
{
    if (p >= total) return false;
    i = (int)(p / rows);            // column (outer loop, main.cc:408)
    j = (int)(p - (long long)i * rows);   // row
    const double2 f = reinterpret_cast<const double2 *>(flow_img)[(size_t)j * cols + i];
    dx = f.x; dy = f.y;
    const double norm = dx * dx + dy * dy;
    return norm > thr;
}

__global__ void __launch_bounds__(kThreads) k_flat_count(const double *__restrict__ flow_img, int rows, int cols,
                                                         double thr, int *block_counts)
{
    const long long total = (long long)rows * cols;
    int cnt = 0;
    for (int s = 0; s < 4; ++s) {
        const long long p = (long long)blockIdx.x * kChunk + s * kThreads + threadIdx.x;
        double dx, dy; int i, j;
        cnt += flow_kept(flow_img, rows, cols, p, total, thr, dx, dy, i, j) ? 1 : 0;
    }
    __shared__ int sh[kWarps];
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w2 = 0; w2 < kWarps; ++w2) t += sh[w2];
        block_counts[blockIdx.x] = t;
    }
}

// Exclusive scan of nb block counts by one CTA of 1024 threads; offsets[nb] = grand total.
__global__ void __launch_bounds__(1024) k_scan_counts(const int *__restrict__ counts, int nb, int *offsets)
{
    __shared__ int seg[1024];
    const int per = (nb + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(nb, lo + per);
    int s = 0;
    for (int b = lo; b < hi; ++b) s += counts[b];
    seg[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int t = 0; t < 1024; ++t) { const int c = seg[t]; seg[t] = run; run += c; }
        offsets[nb] = run;
    }
    __syncthreads();
    int run = seg[threadIdx.x];
    for (int b = lo; b < hi; ++b) { offsets[b] = run; run += counts[b]; }
}

__global__ void __launch_bounds__(kThreads) k_flat_scatter(const double *__restrict__ flow_img, int rows, int cols,
                                                           double thr, double fx, double fy, double cx, double cy,
                                                           double gamma, const int *__restrict__ offsets, double2 *coord,
                                                           double2 *flow, double2 *coord_px, double2 *flow_px,
                                                           int32_t *pixel_index)
{
    const long long total = (long long)rows * cols;
    int base = offsets[blockIdx.x];
    for (int s = 0; s < 4; ++s) {
        const long long p = (long long)blockIdx.x * kChunk + s * kThreads + threadIdx.x;
        double dx = 0, dy = 0; int i = 0, j = 0;
        const bool keep = flow_kept(flow_img, rows, cols, p, total, thr, dx, dy, i, j);
        int tot;
        const int rank = block_rank(keep, tot);
        if (keep) {
            const int pos = base + rank;
            coord_px[pos] = make_double2((double)i, (double)j);
            flow_px[pos] = make_double2(dx, dy);
            flow[pos] = make_double2(dx * gamma / fx, dy * gamma / fy);           // main.cc:424-425
            coord[pos] = make_double2((i - cx) * 1.0 / fx, (j - cy) * 1.0 / fy);  // main.cc:426-427
            if (pixel_index) pixel_index[pos] = (int32_t)p;
        }
        base += tot;
    }
}

// padded tail exactly like the reference constructs the arrays (Ones / Zero, main.cc:401-404)
__global__ void k_flat_tail(const int *__restrict__ offsets, int nb, long long total, double2 *coord, double2 *flow,
                            double2 *coord_px, double2 *flow_px, int32_t *pixel_index)
{
    const int n = offsets[nb];
    for (long long p = n + (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        coord[p] = make_double2(1.0, 1.0);
        coord_px[p] = make_double2(1.0, 1.0);
        flow[p] = make_double2(0.0, 0.0);
        flow_px[p] = make_double2(0.0, 0.0);
        if (pixel_index) pixel_index[p] = -1;
    }
}

int flatten_device(rsdsfm_ctx *ctx, const double *flow_img, int rows, int cols, const double *K4, double gamma,
                   double thr, double *coord, double *flow, double *coord_px, double *flow_px, int32_t *pixel_index,
                   int *n_out)
{
    const long long total = (long long)rows * cols;
    const int nb = (int)((total + kChunk - 1) / kChunk);
    RS_TRY(ensure(ctx, ctx->scan, sizeof(int) * (2 * (size_t)nb + 2)));
    int *counts = (int *)ctx->scan.p, *offsets = counts + nb;
    k_flat_count<<<nb, kThreads, 0, ctx->stream>>>(flow_img, rows, cols, thr, counts);
    k_scan_counts<<<1, 1024, 0, ctx->stream>>>(counts, nb, offsets);
    k_flat_scatter<<<nb, kThreads, 0, ctx->stream>>>(flow_img, rows, cols, thr, K4[0], K4[1], K4[2], K4[3], gamma, offsets,
                                                     (double2 *)coord, (double2 *)flow, (double2 *)coord_px,
                                                     (double2 *)flow_px, pixel_index);
    k_flat_tail<<<grid_for(ctx, total, 4), kThreads, 0, ctx->stream>>>(offsets, nb, total, (double2 *)coord, (double2 *)flow,
                                                                       (double2 *)coord_px, (double2 *)flow_px, pixel_index);
    ctx->launches += 4;
    RS_TRY(ensure_pinned(ctx, 1024));
    RS_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, offsets + nb, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = *(int *)ctx->pinned;
    return RSDSFM_OK;
}

// ---------------------------------------------------------------- a3 / a4
__global__ void k_alpha(const double2 *__restrict__ flow_px, const double2 *__restrict__ q_px, int n, double h,
                        double gamma, double *alpha, double *alpha_k)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double fy = flow_px[i].y;
        if (alpha) alpha[i] = 1 + gamma * fy / h;                                   // minimal.cc:183
        if (alpha_k) {
            const double qy = q_px[i].y;
            const double part1 = gamma * qy / h;                                    // minimal.cc:192-194
            const double part2 = 1.0 + gamma * (qy + fy) / h;
            alpha_k[i] = 0.5 * (part2 * part2 - part1 * part1);
        }
    }
}

int alpha_device(rsdsfm_ctx *ctx, const double *flow_px, const double *q_px, int n, double h, double gamma,
                 double *alpha, double *alpha_k)
{
    if (n <= 0) return RSDSFM_OK;
    k_alpha<<<grid_for(ctx, n, 4), kThreads, 0, ctx->stream>>>((const double2 *)flow_px, (const double2 *)q_px, n, h, gamma,
                                                               alpha, alpha_k);
    ctx->launches++;
    return RSDSFM_OK;
}

// ---------------------------------------------------------------- consensus-set gather
__global__ void __launch_bounds__(kThreads) k_mask_count(const uint8_t *__restrict__ mask, int n, int *block_counts)
{
    int cnt = 0;
    for (int s = 0; s < 4; ++s) {
        const long long p = (long long)blockIdx.x * kChunk + s * kThreads + threadIdx.x;
        cnt += (p < n && mask[p]) ? 1 : 0;
    }
    __shared__ int sh[kWarps];
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w2 = 0; w2 < kWarps; ++w2) t += sh[w2];
        block_counts[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kThreads) k_mask_scatter(const uint8_t *__restrict__ mask, int n,
                                                           const double2 *__restrict__ q, const double *__restrict__ alpha,
                                                           const double *__restrict__ alpha_k,
                                                           const double *__restrict__ inv_depth,
                                                           const int *__restrict__ offsets, double *inliers3,
                                                           double *alpha_in, double *alpha_k_in, int32_t *index_in)
{
    int base = offsets[blockIdx.x];
    for (int s = 0; s < 4; ++s) {
        const long long p = (long long)blockIdx.x * kChunk + s * kThreads + threadIdx.x;
        const bool keep = (p < n) && mask[p];
        int tot;
        const int rank = block_rank(keep, tot);
        if (keep) {
            const size_t pos = (size_t)(base + rank);
            const double2 qq = q[p];
            inliers3[3 * pos] = qq.x;
            inliers3[3 * pos + 1] = qq.y;
            inliers3[3 * pos + 2] = 1.0 / inv_depth[p];                 // minimal.cc:299
            alpha_in[pos] = alpha[p];
            alpha_k_in[pos] = alpha_k[p];
            if (index_in) index_in[pos] = (int32_t)p;
        }
        base += tot;
    }
}

int gather_inliers_device(rsdsfm_ctx *ctx, const double *q, const double *alpha, const double *alpha_k, int n,
                          const uint8_t *mask, const double *inv_depth, double *inliers3, double *alpha_in,
                          double *alpha_k_in, int32_t *index_in, int *m_out)
{
    if (n <= 0) { *m_out = 0; return RSDSFM_OK; }
    const int nb = (n + kChunk - 1) / kChunk;
    RS_TRY(ensure(ctx, ctx->scan, sizeof(int) * (2 * (size_t)nb + 2)));
    int *counts = (int *)ctx->scan.p, *offsets = counts + nb;
    k_mask_count<<<nb, kThreads, 0, ctx->stream>>>(mask, n, counts);
    k_scan_counts<<<1, 1024, 0, ctx->stream>>>(counts, nb, offsets);
    k_mask_scatter<<<nb, kThreads, 0, ctx->stream>>>(mask, n, (const double2 *)q, alpha, alpha_k, inv_depth, offsets,
                                                     inliers3, alpha_in, alpha_k_in, index_in);
    ctx->launches += 3;
    RS_TRY(ensure_pinned(ctx, 1024));
    RS_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, offsets + nb, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *m_out = *(int *)ctx->pinned;
    return RSDSFM_OK;
}

// ---------------------------------------------------------------- compact step inputs -> solver layout
// The "refine + rectify" step fed with what minimal::ransac returned (the winner's consensus mask and inverse
// depths over the flattened points) and the cached flow field itself.  Everything else the reference passes
// around between main.cc:398 and :457 -- normalised coordinates, gamma-scaled flow, alpha / alpha_k, the
// consensus set as (x, y, z) triples -- is a function of those and is rebuilt here, straight into the LM
// solver's tile-blocked layout (lm_layout.h), with the operation order of k_flat_scatter / k_alpha /
// k_mask_scatter (bit-identical values).  Residual block i takes the coordinates and alpha factors of the
// i-th inlier and the flow of flattened POINT i (the reference's pairing, nonlinearRefinement.cc:209-216, Q1).
// flow_img: the caller's row-major field, or (cols < 0) its transposed copy in flattening order (k_flow_transpose)
template <typename F>
__device__ __forceinline__ bool flow_kept_t(const F *__restrict__ flow_img, int rows, int cols, long long p, long long total,
                                            double thr, double &dx, double &dy, int &i, int &j)
{
    if (p >= total) return false;
    i = (int)(p / rows);
    j = (int)(p - (long long)i * rows);
    const size_t at = (cols > 0 ? (size_t)j * cols + i : (size_t)p) * 2;
    dx = (double)flow_img[at]; dy = (double)flow_img[at + 1];      // float32 flow is widened exactly (camera.cc:262-274)
    const double norm = dx * dx + dy * dy;
    return norm > thr;
}

// The flattening walks the image column by column (main.cc:408), the flow field is stored row by row: a thread per
// flattened point would use 8 (float32) or 16 bytes of every 32-byte sector it touches, in each of the three
// compaction passes.  One tiled transpose (coalesced both ways) puts the field into flattening order first.
template <typename F>
__global__ void __launch_bounds__(256) k_flow_transpose(const F *__restrict__ flow_img, int rows, int cols, F *__restrict__ out)
{
    struct P2 { F x, y; };
    __shared__ P2 tile[32][33];
    const P2 *in = reinterpret_cast<const P2 *>(flow_img);
    P2 *o = reinterpret_cast<P2 *>(out);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8 threads
    const int i0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
    for (int r = ty; r < 32; r += 8) {
        const int j = j0 + r, i = i0 + tx;
        if (j < rows && i < cols) tile[r][tx] = in[(size_t)j * cols + i];
    }
    __syncthreads();
    for (int c = ty; c < 32; c += 8) {
        const int i = i0 + c, j = j0 + tx;
        if (j < rows && i < cols) o[(size_t)i * rows + j] = tile[tx][c];
    }
}

// sum of counts[0 .. blockIdx.x)  (whole CTA; every thread receives it): the CTA's offset into the kept points / the
// inliers.  A couple of thousand values out of L2 per CTA -- cheaper than a scan launch between the passes.
__device__ __forceinline__ int preceding_sum(const int *__restrict__ counts)
{
    __shared__ int part[kWarps];
    int v = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += kThreads) v += counts[b];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
    for (int w2 = 0; w2 < kWarps; ++w2) t += part[w2];
    return t;
}

template <typename F>
__global__ void __launch_bounds__(kThreads) k_compact_count_kept(const F *__restrict__ flow_img, int rows, int cols, double thr,
                                                                 int *block_counts)
{
    const long long total = (long long)rows * (cols < 0 ? -cols : cols);      // cols < 0: flow_img is the transposed copy
    int cnt = 0;
    for (int s = 0; s < 4; ++s) {
        const long long p = (long long)blockIdx.x * kChunk + s * kThreads + threadIdx.x;
        double dx, dy; int i, j;
        cnt += flow_kept_t(flow_img, rows, cols, p, total, thr, dx, dy, i, j) ? 1 : 0;
    }
    __shared__ int sh[kWarps];
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w2 = 0; w2 < kWarps; ++w2) t += sh[w2];
        block_counts[blockIdx.x] = t;
    }
}

// second level: inliers among the kept pixels of each chunk (mask is indexed by flattened point)
template <typename F>
__global__ void __launch_bounds__(kThreads) k_compact_count_inliers(const F *__restrict__ flow_img, int rows, int cols, double thr,
                                                                    const int *__restrict__ kept_counts,
                                                                    const uint8_t *__restrict__ mask, int n, int *block_counts)
{
    const long long total = (long long)rows * (cols < 0 ? -cols : cols);      // cols < 0: flow_img is the transposed copy
    int base = preceding_sum(kept_counts);
    int cnt = 0;
    for (int s = 0; s < 4; ++s) {
        const long long p = (long long)blockIdx.x * kChunk + s * kThreads + threadIdx.x;
        double dx, dy; int i, j;
        const bool keep = flow_kept_t(flow_img, rows, cols, p, total, thr, dx, dy, i, j);
        int tot;
        const int point = base + block_rank(keep, tot);
        cnt += (keep && point < n && mask[point]) ? 1 : 0;
        base += tot;
    }
    __shared__ int sh[kWarps];
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w2 = 0; w2 < kWarps; ++w2) t += sh[w2];
        block_counts[blockIdx.x] = t;
    }
}

template <typename F>
__global__ void __launch_bounds__(kThreads) k_compact_scatter(const F *__restrict__ flow_img, int rows, int cols, double thr,
                                                              double fx, double fy, double cx, double cy, double gamma,
                                                              const int *__restrict__ kept_counts,
                                                              const int *__restrict__ inl_counts,
                                                              const uint8_t *__restrict__ mask,
                                                              const double *__restrict__ inv_depth, int n, int m, double2 *blk,
                                                              double *d0, double *z_in, double2 *xy, int *input_flag)
{
    const long long total = (long long)rows * (cols < 0 ? -cols : cols);      // cols < 0: flow_img is the transposed copy
    int base = preceding_sum(kept_counts), ibase = preceding_sum(inl_counts);
    const double h = (double)rows;
    int bad = 0;
    for (int s = 0; s < 4; ++s) {
        const long long p = (long long)blockIdx.x * kChunk + s * kThreads + threadIdx.x;
        double dx = 0, dy = 0; int i = 0, j = 0;
        const bool keep = flow_kept_t(flow_img, rows, cols, p, total, thr, dx, dy, i, j);
        int tot, itot;
        const int point = base + block_rank(keep, tot);
        const bool inl = keep && point < n && mask[point];
        const int pos = ibase + block_rank(inl, itot);
        if (keep && point < m)                                               // flow of flattened point `point` -> residual block `point`
            blk[blk_index(point, 1)] = make_double2(dx * gamma / fx, dy * gamma / fy);          // main.cc:424-425
        if (inl && pos < m) {
            const double2 q = make_double2((i - cx) * 1.0 / fx, (j - cy) * 1.0 / fy);           // main.cc:426-427
            blk[blk_index(pos, 0)] = q;
            xy[pos] = q;
            const double qy = (double)j;
            const double alpha = 1 + gamma * dy / h;                                            // minimal.cc:183
            const double part1 = gamma * qy / h;                                                // minimal.cc:192-194
            const double part2 = 1.0 + gamma * (qy + dy) / h;
            blk[blk_index(pos, 2)] = make_double2(alpha, 0.5 * (part2 * part2 - part1 * part1));
            const double z = 1.0 / inv_depth[point];                                            // minimal.cc:299
            z_in[pos] = z;
            const double d = 1.0 / z;                                                           // nonlinearRefinement.cc:213
            d0[pos] = d;
            if (!isfinite(d)) bad = 1;
        }
        base += tot; ibase += itot;
    }
    // the caller's n / m must be what the flow field and the mask really hold (the last CTA ends on the totals)
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0 && (base != n || ibase != m)) atomicOr(input_flag, 2);
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(input_flag, 1);
}

template <typename F>
static int compact_build_t(rsdsfm_ctx *ctx, const F *flow_img, int rows, int cols, const double *K4, double gamma, double thr,
                           const uint8_t *mask, const double *inv_depth, int n, int m, double2 *blk, double *d0, double *z_in,
                           double2 *xy, int *input_flag)
{
    const long long total = (long long)rows * cols;
    const int nb = (int)((total + kChunk - 1) / kChunk);
    RS_TRY(ensure(ctx, ctx->scan, sizeof(int) * (2 * (size_t)nb + 4)));
    int *counts = (int *)ctx->scan.p, *icounts = counts + nb;
    RS_TRY(ensure(ctx, ctx->flow_t, sizeof(F) * 2 * (size_t)total));
    F *ft = (F *)ctx->flow_t.p;
    k_flow_transpose<F><<<dim3((cols + 31) / 32, (rows + 31) / 32), 256, 0, ctx->stream>>>(flow_img, rows, cols, ft);
    k_compact_count_kept<F><<<nb, kThreads, 0, ctx->stream>>>(ft, rows, -cols, thr, counts);
    k_compact_count_inliers<F><<<nb, kThreads, 0, ctx->stream>>>(ft, rows, -cols, thr, counts, mask, n, icounts);
    k_compact_scatter<F><<<nb, kThreads, 0, ctx->stream>>>(ft, rows, -cols, thr, K4[0], K4[1], K4[2], K4[3], gamma, counts, icounts,
                                                          mask, inv_depth, n, m, blk, d0, z_in, xy, input_flag);
    ctx->launches += 4;
    RS_CUDA(ctx, cudaGetLastError());
    return RSDSFM_OK;
}

int compact_build_device(rsdsfm_ctx *ctx, const void *flow_img, int flow_f32, int rows, int cols, const double *K4, double gamma,
                         double thr, const uint8_t *mask, const double *inv_depth, int n, int m, void *blk, double *d0,
                         double *z_in, double *xy, int *input_flag)
{
    if (flow_f32)
        return compact_build_t<float>(ctx, (const float *)flow_img, rows, cols, K4, gamma, thr, mask, inv_depth, n, m, (double2 *)blk,
                                      d0, z_in, (double2 *)xy, input_flag);
    return compact_build_t<double>(ctx, (const double *)flow_img, rows, cols, K4, gamma, thr, mask, inv_depth, n, m, (double2 *)blk, d0,
                                   z_in, (double2 *)xy, input_flag);
}

}  // namespace rsdsfm
