// ransac.cu -- Minimal-stage RANSAC scoring of a list of RS differential-epipolar hypotheses over
// all flow correspondences (minimal::ransac, minimal.cc:209-306), including the per-hypothesis
// inverse-depth estimation the reference runs through Ceres for every trial
// (nonlinear_refinement::estimateInverseDepths, nonlinearRefinement.cc:109-180).
//
// Compiled with --fmad=false.  Inlier sets have to be bit-exact for a given hypothesis list, so
// the per-point arithmetic below is the plain IEEE-double operation sequence of the reference's
// residual functor (nonlinearRefinement.cc:32-52), of the Schur-eliminated LM step Ceres takes
// on an all-e-block problem and of the scoring loop (minimal.cc:255-275).
//
// The depth problem has no f-blocks: each LM iteration is a per-point damped update that depends
// on the rest of the image only through the trust-region radius and the accept/terminate
// decisions, which are functions of a few global sums.  Per (point, hypothesis) only the inverse
// depth is kept between passes, in two ping-pong planes (current point / candidate); a pass
// evaluates the current point, takes the candidate step, writes the candidate depth and reduces
// the sums the controller needs; the host runs the Ceres logic per hypothesis and flips the planes
// of the hypotheses whose step was accepted.  The kernel is bound by FP64 issue (IEEE divisions and
// square roots that bit-exactness forbids replacing): algorithmic traffic is 48 B per point per
// hypothesis (L2-resident: the point arrays are 100 MB) plus 16 B of depth planes per pass.
#include "common.cuh"
#include "lm_controller.h"
#include "solve9.h"

namespace rsdsfm {

// Hypotheses per CTA row (blockIdx.x) and CTAs per SM.  Measured at 1080p, H = 16 (tools/stage_times.py):
// (4,1) 3.23 ms, (4,2) 2.90, (2,2) 2.53, (2,3) 2.45, (1,3) 2.31, (1,4) 2.20 ms -- the IEEE division /
// square-root sequences are long dependent chains, so occupancy (64 registers, 32 warps/SM) beats
// sharing one point load between several hypotheses; the point arrays (100 MB) stay L2-resident.
constexpr int kHG = 1;
constexpr int kPassOcc = 4, kScoreOcc = 6;
constexpr int kHChunk = 64;     // hypotheses scored per batch (bounds the depth planes: 16 B x n x kHChunk)

struct HypDev {
    double w[3], v[3], k;
    double c2;                  // 2 / (2 + k)
    double radius;              // radius of the candidate step of this pass
    int n_acc;                  // accepted steps so far (0: the current point is d = 1 everywhere)
    int active;                 // still iterating
    int failed;                 // solver FAILURE: depths keep their start value 1.0
    int cur;                    // which depth plane holds the current point
};

enum { RS_COST = 0, RS_SUMSQ_D = 1, RS_MCC = 2, RS_STEP_SQ = 3, RS_CAND_COST = 4, RS_NS = 5 };
enum { RM_GMAX_E = 0, RM_BAD = 1, RM_BAD_STEP = 2, RM_BAD_CAND = 3, RM_NM = 4 };

// Per (point, hypothesis) constants of the depth problem.
struct DepthTerms {
    double nb, g0, g1, t00, t01, t02, t10, t11, t12;   // residual: r = u - nb * ((d*g + t0) -/+ ...)
    double ux, uy;
    double E0, E1, scale, e0, e1, ee, diag;
};

__device__ __forceinline__ void depth_terms(double x, double y, double ux, double uy, double alpha, double alpha_k,
                                            const HypDev &h, double min_diag, double max_diag, DepthTerms &T)
{
    const double beta = h.c2 * (alpha + h.k * alpha_k);
    T.nb = beta * -1.0;
    T.g0 = x * h.v[2] - h.v[0];
    T.g1 = y * h.v[2] - h.v[1];
    const double xy = x * y;
    T.t00 = xy * h.w[0];  T.t01 = (1.0 + x * x) * h.w[1];  T.t02 = y * h.w[2];
    T.t10 = (1.0 + y * y) * h.w[0];  T.t11 = xy * h.w[1];  T.t12 = x * h.w[2];
    T.ux = ux; T.uy = uy;
    T.E0 = beta * T.g0;
    T.E1 = beta * T.g1;
    T.scale = 1.0 / (1.0 + sqrt(T.E0 * T.E0 + T.E1 * T.E1));      // Jacobi scaling, iteration 0
    T.e0 = T.E0 * T.scale;
    T.e1 = T.E1 * T.scale;
    T.ee = T.e0 * T.e0 + T.e1 * T.e1;
    T.diag = fmin(fmax(T.ee, min_diag), max_diag);
}

__device__ __forceinline__ void depth_residual(const DepthTerms &T, double d, double &r0, double &r1)
{
    r0 = T.ux - T.nb * (d * T.g0 + T.t00 - T.t01 + T.t02);
    r1 = T.uy - T.nb * (d * T.g1 + T.t10 - T.t11 - T.t12);
}

// One LM step of the 1x1 e-block with trust-region radius R: returns the scaled step (-y_e).
__device__ __forceinline__ double depth_step(const DepthTerms &T, double r0, double r1, double R)
{
    const double De = sqrt(T.diag / R);
    const double ete = T.ee + De * De;
    const double inv = 1.0 / ete;
    const double ye = (T.e0 * r0 + T.e1 * r1) * inv;
    return -ye;
}

// 1.0 if any argument is NaN or +-Inf, else 0.0 (x - x is 0 for finite x and NaN otherwise)
__device__ __forceinline__ double any_nonfinite4(double a, double b, double c, double d)
{
    const double z = ((a - a) + (b - b)) + ((c - c) + (d - d));
    return (z == 0.0) ? 0.0 : 1.0;
}

// depth plane `which` (0/1) of hypothesis slot `hs` (index inside the batch)
__device__ __forceinline__ size_t plane(int hs, int which, int n) { return ((size_t)hs * 2 + (size_t)which) * (size_t)n; }

template <int HG>
__device__ __forceinline__ void load_hyps(HypDev *sh, const HypDev *hyps, int h0, int H)
{
    for (int t = threadIdx.x; t < (int)(HG * sizeof(HypDev) / sizeof(int)); t += blockDim.x) {
        const int hh = t / (int)(sizeof(HypDev) / sizeof(int));
        int val = 0;
        if (h0 + hh < H) val = reinterpret_cast<const int *>(hyps + h0)[t];
        reinterpret_cast<int *>(sh)[t] = val;
    }
    __syncthreads();
}

template <int HG, int OCC>
__global__ void __launch_bounds__(kThreads, OCC) k_ransac_pass(const double2 *__restrict__ q, const double2 *__restrict__ u,
                                                          const double *__restrict__ alpha,
                                                          const double *__restrict__ alpha_k, int n,
                                                          const HypDev *__restrict__ hyps, int H, double min_diag,
                                                          double max_diag, double *__restrict__ depth,
                                                          double *__restrict__ partials)
{
    __shared__ HypDev sh[HG];
    // grid = (hypothesis rows, point chunks): consecutive CTAs score DIFFERENT hypotheses on the SAME points,
    // so the point arrays are fetched from DRAM once per pass and served from L2 to the other hypotheses
    const int h0 = blockIdx.x * HG, chunk = blockIdx.y, nchunks = gridDim.y;
    load_hyps<HG>(sh, hyps, h0, H);
    double s[HG * RS_NS], mx[HG * RM_NM];
#pragma unroll
    for (int j = 0; j < HG * RS_NS; ++j) s[j] = 0.0;
#pragma unroll
    for (int j = 0; j < HG * RM_NM; ++j) mx[j] = 0.0;

    for (int i = chunk * blockDim.x + threadIdx.x; i < n; i += nchunks * blockDim.x) {
        const double2 qq = q[i], uu = u[i];
        const double al = alpha[i], alk = alpha_k[i];
#pragma unroll
        for (int g = 0; g < HG; ++g) {
            const HypDev &h = sh[g];
            if (!h.active) continue;
            DepthTerms T;
            depth_terms(qq.x, qq.y, uu.x, uu.y, al, alk, h, min_diag, max_diag, T);
            const double d = (h.n_acc > 0) ? depth[plane(h0 + g, h.cur, n) + i] : 1.0;     // nonlinearRefinement.cc:140
            double r0, r1;
            depth_residual(T, d, r0, r1);
            // evaluation at x
            s[g * RS_NS + RS_COST] += 0.5 * (r0 * r0 + r1 * r1);
            s[g * RS_NS + RS_SUMSQ_D] += d * d;
            const double ge = T.E0 * r0 + T.E1 * r1;
            const double proj = d + (-ge);
            mx[g * RM_NM + RM_GMAX_E] = fmax(mx[g * RM_NM + RM_GMAX_E], fabs(d - proj));
            mx[g * RM_NM + RM_BAD] = fmax(mx[g * RM_NM + RM_BAD], any_nonfinite4(r0, r1, T.E0, T.E1));
            // candidate step
            const double step_e = depth_step(T, r0, r1, h.radius);
            const double mr0 = T.e0 * step_e, mr1 = T.e1 * step_e;
            s[g * RS_NS + RS_MCC] += mr0 * (r0 + mr0 / 2.0) + mr1 * (r1 + mr1 / 2.0);
            const double dc = d + step_e * T.scale;
            depth[plane(h0 + g, h.cur ^ 1, n) + i] = dc;
            const double dd = d - dc;
            s[g * RS_NS + RS_STEP_SQ] += dd * dd;
            double c0, c1;
            depth_residual(T, dc, c0, c1);
            s[g * RS_NS + RS_CAND_COST] += 0.5 * (c0 * c0 + c1 * c1);
            mx[g * RM_NM + RM_BAD_STEP] = fmax(mx[g * RM_NM + RM_BAD_STEP], bad_flag(step_e));
            mx[g * RM_NM + RM_BAD_CAND] = fmax(mx[g * RM_NM + RM_BAD_CAND], any_nonfinite4(c0, c1, 0.0, 0.0));
        }
    }
    block_reduce_store<HG * RS_NS, HG * RM_NM>(s, mx, partials + (size_t)blockIdx.x * nchunks * (HG * (RS_NS + RM_NM)), chunk);
}

// Scoring loop minimal.cc:255-275 at the final depths: per hypothesis inlier count and error sum.
__device__ __forceinline__ double ransac_error(double x, double y, double ux, double uy, double alpha, double alpha_k,
                                               const HypDev &h, double d)
{
    const double av0 = 1.0 * h.v[0] + 0.0 * h.v[1] + (-x) * h.v[2];
    const double av1 = 0.0 * h.v[0] + 1.0 * h.v[1] + (-y) * h.v[2];
    const double bw0 = (-x * y) * h.w[0] + (1 + x * x) * h.w[1] + (-y) * h.w[2];
    const double bw1 = (-(1 + y * y)) * h.w[0] + (x * y) * h.w[1] + x * h.w[2];
    const double beta = (alpha + h.k * alpha_k) * h.c2;
    const double ue0 = beta * (av0 * d + bw0);
    const double ue1 = beta * (av1 * d + bw1);
    const double dx = ue0 - ux, dy = ue1 - uy;
    return sqrt(dx * dx + dy * dy);
}

__device__ __forceinline__ double final_depth(const HypDev &h, const double *depth, int hs, int n, int i)
{
    return (h.failed || h.n_acc == 0) ? 1.0 : depth[plane(hs, h.cur, n) + i];
}

template <int HG, int OCC>
__global__ void __launch_bounds__(kThreads, OCC) k_ransac_score(const double2 *__restrict__ q, const double2 *__restrict__ u,
                                                           const double *__restrict__ alpha,
                                                           const double *__restrict__ alpha_k, int n,
                                                           const HypDev *__restrict__ hyps, int H, double tol,
                                                           const double *__restrict__ depth, double *__restrict__ partials)
{
    __shared__ HypDev sh[HG];
    const int h0 = blockIdx.x * HG, chunk = blockIdx.y, nchunks = gridDim.y;
    load_hyps<HG>(sh, hyps, h0, H);
    double s[HG * 2], mx[1] = {0.0};
#pragma unroll
    for (int j = 0; j < HG * 2; ++j) s[j] = 0.0;
    for (int i = chunk * blockDim.x + threadIdx.x; i < n; i += nchunks * blockDim.x) {
        const double2 qq = q[i], uu = u[i];
        const double al = alpha[i], alk = alpha_k[i];
#pragma unroll
        for (int g = 0; g < HG; ++g) {
            if (h0 + g >= H) continue;
            const HypDev &h = sh[g];
            const double d = final_depth(h, depth, h0 + g, n, i);
            const double err = ransac_error(qq.x, qq.y, uu.x, uu.y, al, alk, h, d);
            if (err < tol) { s[g * 2] += 1.0; s[g * 2 + 1] += err; }
        }
    }
    block_reduce_store<HG * 2, 0>(s, mx, partials + (size_t)blockIdx.x * nchunks * (HG * 2), chunk);
}

__global__ void __launch_bounds__(kThreads) k_ransac_winner(const double2 *__restrict__ q, const double2 *__restrict__ u,
                                                            const double *__restrict__ alpha,
                                                            const double *__restrict__ alpha_k, int n,
                                                            const HypDev *__restrict__ hyps, int best, double tol,
                                                            const double *__restrict__ depth, uint8_t *mask, double *inv_depth)
{
    __shared__ HypDev h;
    for (int t = threadIdx.x; t < (int)(sizeof(HypDev) / sizeof(int)); t += blockDim.x)
        reinterpret_cast<int *>(&h)[t] = reinterpret_cast<const int *>(hyps + best)[t];
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double2 qq = q[i], uu = u[i];
        const double d = final_depth(h, depth, best, n, i);
        const double err = ransac_error(qq.x, qq.y, uu.x, uu.y, alpha[i], alpha_k[i], h, d);
        if (inv_depth) inv_depth[i] = d;
        if (mask) mask[i] = (err < tol) ? 1 : 0;
    }
}

// rows of `width` doubles per CTA, grid (gx, gy): out[by*width + j] = sum/max over bx.  One warp per
// column: lane l combines rows l, l+32, ... in ascending order, then a fixed shuffle tree.
__global__ void __launch_bounds__(kThreads) k_ransac_reduce(const double *__restrict__ partials, int gx, int width, int nsum,
                                                            double *out)
{
    const int by = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = warp; j < width; j += kWarps) {
        const double *p = partials + (size_t)by * gx * width + j;
        const bool is_sum = j < nsum;
        double v = 0.0;                                       // maxima are all >= 0
        for (int b = lane; b < gx; b += 32) { const double x = p[(size_t)b * width]; v = is_sum ? v + x : fmax(v, x); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double x = __shfl_xor_sync(0xffffffffu, v, o);
            v = is_sum ? v + x : fmax(v, x);
        }
        if (lane == 0) out[(size_t)by * width + j] = v;
    }
}

__global__ void k_gather_samples(const double2 *__restrict__ q, const double2 *__restrict__ u,
                                 const double *__restrict__ alpha, const double *__restrict__ alpha_k,
                                 const int32_t *__restrict__ samples, int count, int n, double *out6)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int idx = samples[t];
    if (idx < 0 || idx >= n) { for (int j = 0; j < 6; ++j) out6[6 * t + j] = NAN; return; }
    out6[6 * t + 0] = q[idx].x; out6[6 * t + 1] = q[idx].y;
    out6[6 * t + 2] = u[idx].x; out6[6 * t + 3] = u[idx].y;
    out6[6 * t + 4] = alpha[idx]; out6[6 * t + 5] = alpha_k[idx];
}

// Scores one batch of Hb <= kHChunk hypotheses (host array hyps7) on device-resident points.
static int score_batch(rsdsfm_ctx *ctx, const double *q, const double *u, const double *alpha, const double *alpha_k, int n,
                       const double *hyps7, int Hb, double tol, int *counts, double *sumerr, HypDev **hd_out)
{
    rsdsfm_lm_options opt;
    rsdsfm_lm_default_options(&opt);
    constexpr int hg = kHG;
    const int gy = (Hb + hg - 1) / hg;
    const int gx = grid_for(ctx, n, 2);
    const int widthP = hg * (RS_NS + RM_NM), widthS = hg * 2;
    RS_TRY(ensure(ctx, ctx->partials, sizeof(double) * (size_t)gx * gy * widthP));
    RS_TRY(ensure(ctx, ctx->sums, sizeof(double) * (size_t)gy * widthP));
    RS_TRY(ensure(ctx, ctx->hyp, sizeof(HypDev) * (size_t)gy * hg));
    RS_TRY(ensure(ctx, ctx->rdepth, sizeof(double) * 2 * (size_t)gy * hg * (size_t)(n > 0 ? n : 1)));
    RS_TRY(ensure_pinned(ctx, sizeof(HypDev) * (size_t)gy * hg + sizeof(double) * (size_t)gy * widthP + 64));
    HypDev *hh = (HypDev *)ctx->pinned;
    double *hs = (double *)((char *)ctx->pinned + sizeof(HypDev) * (size_t)gy * hg);
    HypDev *hd = (HypDev *)ctx->hyp.p;
    double *partials = (double *)ctx->partials.p, *sums = (double *)ctx->sums.p, *depth = (double *)ctx->rdepth.p;
    *hd_out = hd;

    std::vector<LmController> ctl((size_t)Hb);
    std::vector<LmNext> pend((size_t)Hb, LM_RUN_A);
    memset(hh, 0, sizeof(HypDev) * (size_t)gy * hg);
    int n_active = 0;
    for (int h = 0; h < Hb; ++h) {
        const double *p = hyps7 + 7 * h;
        for (int j = 0; j < 3; ++j) { hh[h].w[j] = p[j]; hh[h].v[j] = p[3 + j]; }
        hh[h].k = p[6];
        hh[h].c2 = 2.0 / (2.0 + p[6]);
        ctl[h].init(opt, 0, nullptr);
        hh[h].radius = ctl[h].radius;
        bool finite = true;
        for (int j = 0; j < 7; ++j) finite = finite && isfinite(p[j]);
        // solver.cc: non-finite parameter values => FAILURE, depths stay at 1.0.  n == 0: nothing to do.
        hh[h].active = (finite && n > 0) ? 1 : 0;
        hh[h].failed = finite ? 0 : 1;
        n_active += hh[h].active;
    }
    const dim3 grid(gy, gx);          // x = hypothesis row (fastest), y = point chunk
    while (n_active > 0) {
        RS_CUDA(ctx, cudaMemcpyAsync(hd, hh, sizeof(HypDev) * (size_t)gy * hg, cudaMemcpyHostToDevice, ctx->stream));
        k_ransac_pass<kHG, kPassOcc><<<grid, kThreads, 0, ctx->stream>>>((const double2 *)q, (const double2 *)u, alpha, alpha_k, n, hd, Hb,
                                                                          opt.min_lm_diagonal, opt.max_lm_diagonal, depth, partials);
        k_ransac_reduce<<<gy, kThreads, 0, ctx->stream>>>(partials, gx, widthP, hg * RS_NS, sums);
        ctx->launches += 2;
        RS_CUDA(ctx, cudaMemcpyAsync(hs, sums, sizeof(double) * (size_t)gy * widthP, cudaMemcpyDeviceToHost, ctx->stream));
        RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        n_active = 0;
        for (int h = 0; h < Hb; ++h) {
            if (!hh[h].active) continue;
            const double *row = hs + (size_t)(h / hg) * widthP;
            const double *ss = row + (h % hg) * RS_NS, *mm = row + hg * RS_NS + (h % hg) * RM_NM;
            EvalSums e;
            memset(&e, 0, sizeof e);
            e.cost = ss[RS_COST]; e.sumsq_d = ss[RS_SUMSQ_D]; e.gmax_e = mm[RM_GMAX_E]; e.bad = mm[RM_BAD];
            CandSums c;
            c.mcc = ss[RS_MCC]; c.step_sq = ss[RS_STEP_SQ]; c.cand_cost = ss[RS_CAND_COST];
            c.bad_step = mm[RM_BAD_STEP]; c.bad_cand = mm[RM_BAD_CAND];
            LmNext next = pend[h];
            if (next == LM_RUN_A) next = ctl[h].on_eval(e);               // evaluation at a new point
            while (next == LM_SOLVE) next = ctl[h].solve_step(nullptr);   // no f-blocks: nothing to solve
            if (next == LM_RUN_B) next = ctl[h].on_candidate(c);
            pend[h] = next;
            if (next == LM_DONE) {
                // the candidate of the terminating iteration is not taken (Ceres tests the tolerances first)
                hh[h].active = 0;
                hh[h].failed = (ctl[h].termination == RSDSFM_FAILURE) ? 1 : 0;
            } else {
                if (ctl[h].accepted_last) { hh[h].cur ^= 1; hh[h].n_acc++; }   // the candidate plane becomes the current point
                hh[h].radius = ctl[h].radius;
                n_active++;
            }
        }
    }
    // scoring
    RS_CUDA(ctx, cudaMemcpyAsync(hd, hh, sizeof(HypDev) * (size_t)gy * hg, cudaMemcpyHostToDevice, ctx->stream));
    if (n > 0) {
        k_ransac_score<kHG, kScoreOcc><<<grid, kThreads, 0, ctx->stream>>>((const double2 *)q, (const double2 *)u, alpha, alpha_k, n, hd, Hb,
                                                                            tol, depth, partials);
        k_ransac_reduce<<<gy, kThreads, 0, ctx->stream>>>(partials, gx, widthS, widthS, sums);
        ctx->launches += 2;
        RS_CUDA(ctx, cudaMemcpyAsync(hs, sums, sizeof(double) * (size_t)gy * widthS, cudaMemcpyDeviceToHost, ctx->stream));
        RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int j = 0; j < gy * widthS; ++j) hs[j] = 0.0;
    }
    for (int h = 0; h < Hb; ++h) {
        const double *row = hs + (size_t)(h / hg) * widthS + (h % hg) * 2;
        counts[h] = (int)row[0];
        sumerr[h] = row[1];
    }
    return RSDSFM_OK;
}

// Scores H hypotheses (host array hyps7) on device-resident points, kHChunk at a time.
int ransac_score_device(rsdsfm_ctx *ctx, const double *q, const double *u, const double *alpha, const double *alpha_k,
                        int n, const double *hyps7, int H, double tol, int *counts, double *sumerr, int *best_idx,
                        uint8_t *mask_best, double *inv_depth_best)
{
    *best_idx = -1;
    if (H <= 0) return RSDSFM_OK;
    int best = -1, best_count = -1;
    double best_err = 0.0;
    std::vector<int> cnt((size_t)kHChunk);
    std::vector<double> err((size_t)kHChunk);
    for (int b0 = 0; b0 < H; b0 += kHChunk) {
        const int Hb = (H - b0 < kHChunk) ? (H - b0) : kHChunk;
        HypDev *hd = nullptr;
        RS_TRY(score_batch(ctx, q, u, alpha, alpha_k, n, hyps7 + 7 * (size_t)b0, Hb, tol, cnt.data(), err.data(), &hd));
        int batch_best = -1;
        for (int h = 0; h < Hb; ++h) {
            if (counts) counts[b0 + h] = cnt[(size_t)h];
            if (sumerr) sumerr[b0 + h] = err[(size_t)h];
            if (cnt[(size_t)h] > best_count || (cnt[(size_t)h] == best_count && err[(size_t)h] < best_err)) {   // minimal.cc:278
                best_count = cnt[(size_t)h]; best_err = err[(size_t)h]; best = b0 + h; batch_best = h;
            }
        }
        // the depth planes only live until the next batch: extract the leader's consensus set now
        if (batch_best >= 0 && n > 0 && (mask_best || inv_depth_best)) {
            k_ransac_winner<<<grid_for(ctx, n, 4), kThreads, 0, ctx->stream>>>((const double2 *)q, (const double2 *)u, alpha, alpha_k,
                                                                               n, hd, batch_best, tol, (const double *)ctx->rdepth.p,
                                                                               mask_best, inv_depth_best);
            ctx->launches++;
        }
    }
    *best_idx = best;
    return RSDSFM_OK;
}

// Gathers the sampled correspondences, fits each hypothesis with the 9-point solver (host).
int ransac_fit_device(rsdsfm_ctx *ctx, const double *q, const double *u, const double *alpha, const double *alpha_k,
                      int n, int use_alpha_k, const int32_t *samples_host, int H, double *hyps7_out)
{
    const int count = H * 9;
    RS_TRY(ensure(ctx, ctx->misc, sizeof(double) * 6 * (size_t)count + sizeof(int32_t) * (size_t)count));
    RS_TRY(ensure_pinned(ctx, sizeof(double) * 6 * (size_t)count + 64));
    double *d6 = (double *)ctx->misc.p;
    int32_t *ds = (int32_t *)(d6 + 6 * (size_t)count);
    RS_CUDA(ctx, cudaMemcpyAsync(ds, samples_host, sizeof(int32_t) * (size_t)count, cudaMemcpyHostToDevice, ctx->stream));
    k_gather_samples<<<(count + 127) / 128, 128, 0, ctx->stream>>>((const double2 *)q, (const double2 *)u, alpha, alpha_k, ds,
                                                                  count, n, d6);
    ctx->launches++;
    RS_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, d6, sizeof(double) * 6 * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const double *h6 = (const double *)ctx->pinned;
    for (int h = 0; h < H; ++h) {
        double q9[18], u9[18], a9[9], ak9[9];
        for (int j = 0; j < 9; ++j) {
            const double *p = h6 + 6 * (size_t)(h * 9 + j);
            q9[2 * j] = p[0]; q9[2 * j + 1] = p[1]; u9[2 * j] = p[2]; u9[2 * j + 1] = p[3]; a9[j] = p[4]; ak9[j] = p[5];
        }
        s9::calculate_velocities(q9, u9, a9, ak9, use_alpha_k != 0, hyps7_out + 7 * h);
    }
    return RSDSFM_OK;
}

}  // namespace rsdsfm
