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
// the sums the controller needs; a one-block-per-hypothesis kernel (k_ransac_ctl) runs the Ceres logic and
// flips the planes of the hypotheses whose step was accepted -- no host round trip per pass.  The kernel is bound by FP64 issue (IEEE divisions and
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
    if (HG == 1 && !sh[0].active) return;          // finished (or never started): its row of partials is not read
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
    int cnt[HG];
#pragma unroll
    for (int j = 0; j < HG * 2; ++j) s[j] = 0.0;
#pragma unroll
    for (int g = 0; g < HG; ++g) cnt[g] = 0;
    for (int i = chunk * blockDim.x + threadIdx.x; i < n; i += nchunks * blockDim.x) {
        const double2 qq = q[i], uu = u[i];
        const double al = alpha[i], alk = alpha_k[i];
#pragma unroll
        for (int g = 0; g < HG; ++g) {
            if (h0 + g >= H) continue;
            const HypDev &h = sh[g];
            const double d = final_depth(h, depth, h0 + g, n, i);
            const double err = ransac_error(qq.x, qq.y, uu.x, uu.y, al, alk, h, d);
            const bool inl = err < tol;
            cnt[g] += __popc(__ballot_sync(__activemask(), inl));        // every lane keeps the warp's count
            if (inl) s[g * 2 + 1] += err;
        }
    }
    // lane 0 of every warp carries the warp's ballot count into the block sum (exact: integers below 2^53)
#pragma unroll
    for (int g = 0; g < HG; ++g) s[g * 2] = ((threadIdx.x & 31) == 0) ? (double)cnt[g] : 0.0;
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

// ---- the Ceres logic of every hypothesis' depth-only problem, on the device -------------------------------
// Per batch: HypDev (what the pass kernels read), the LmController of each hypothesis and its pending request.
struct HypCtl {
    LmController ctl;
    int pend;                   // LmNext the next sums answer (LM_RUN_A: evaluation at a new point, LM_RUN_B: candidate)
};

__global__ void k_ransac_init(const double *__restrict__ hyps7, int Hb, int n, rsdsfm_lm_options opt, HypDev *hd, HypCtl *hc,
                              int *n_active)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= Hb) return;
    const double *p = hyps7 + 7 * h;
    HypDev d;
    memset(&d, 0, sizeof d);
    bool finite = true;
    for (int j = 0; j < 7; ++j) finite = finite && isfinite(p[j]);
    for (int j = 0; j < 3; ++j) { d.w[j] = p[j]; d.v[j] = p[3 + j]; }
    d.k = p[6];
    d.c2 = 2.0 / (2.0 + p[6]);
    hc[h].ctl.init(opt, 0, nullptr);
    hc[h].pend = (int)LM_RUN_A;
    d.radius = hc[h].ctl.radius;
    // solver.cc: non-finite parameter values => FAILURE, depths stay at 1.0.  n == 0: nothing to do.
    d.active = (finite && n > 0) ? 1 : 0;
    d.failed = finite ? 0 : 1;
    hd[h] = d;
    if (d.active) atomicAdd(n_active, 1);
}

// One block per hypothesis: the sums of its pass (same order as k_ransac_reduce: one warp per column, lane l
// combines rows l, l+32, ... in ascending order, then a fixed shuffle tree), then the controller step on thread 0.
__global__ void __launch_bounds__(kThreads) k_ransac_ctl(const double *__restrict__ partials, int gx, HypDev *hd, HypCtl *hc, int *n_active)
{
    constexpr int width = RS_NS + RM_NM;
    __shared__ double row[width];
    const int h = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (!hd[h].active) return;
    for (int j = warp; j < width; j += kWarps) {
        const double *p = partials + (size_t)h * gx * width + j;
        const bool is_sum = j < RS_NS;
        double v = 0.0;                                       // maxima are all >= 0
        for (int b = lane; b < gx; b += 32) { const double x = p[(size_t)b * width]; v = is_sum ? v + x : fmax(v, x); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double x = __shfl_xor_sync(0xffffffffu, v, o);
            v = is_sum ? v + x : fmax(v, x);
        }
        if (lane == 0) row[j] = v;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const double *ss = row, *mm = row + RS_NS;
    LmController &c = hc[h].ctl;
    EvalSums e;
    memset(&e, 0, sizeof e);
    e.cost = ss[RS_COST]; e.sumsq_d = ss[RS_SUMSQ_D]; e.gmax_e = mm[RM_GMAX_E]; e.bad = mm[RM_BAD];
    CandSums cs;
    cs.mcc = ss[RS_MCC]; cs.step_sq = ss[RS_STEP_SQ]; cs.cand_cost = ss[RS_CAND_COST];
    cs.bad_step = mm[RM_BAD_STEP]; cs.bad_cand = mm[RM_BAD_CAND];
    LmNext next = (LmNext)hc[h].pend;
    if (next == LM_RUN_A) next = c.on_eval(e);               // evaluation at a new point
    while (next == LM_SOLVE) next = c.solve_step(nullptr);   // no f-blocks: nothing to solve
    if (next == LM_RUN_B) next = c.on_candidate(cs);
    hc[h].pend = (int)next;
    if (next == LM_DONE) {
        // the candidate of the terminating iteration is not taken (Ceres tests the tolerances first)
        hd[h].active = 0;
        hd[h].failed = (c.termination == RSDSFM_FAILURE) ? 1 : 0;
        atomicSub(n_active, 1);
    } else {
        if (c.accepted_last) { hd[h].cur ^= 1; hd[h].n_acc++; }   // the candidate plane becomes the current point
        hd[h].radius = c.radius;
    }
}

// Exact tie-break (minimal.cc:255-278 adds the errors of the inliers in index order): per-point errors of one
// hypothesis (0 for outliers) ...
__global__ void __launch_bounds__(kThreads) k_ransac_errors(const double2 *__restrict__ q, const double2 *__restrict__ u,
                                                            const double *__restrict__ alpha, const double *__restrict__ alpha_k, int n,
                                                            const HypDev *__restrict__ hyps, int which, double tol,
                                                            const double *__restrict__ depth, double *err_out)
{
    const HypDev h = hyps[which];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double2 qq = q[i], uu = u[i];
        const double err = ransac_error(qq.x, qq.y, uu.x, uu.y, alpha[i], alpha_k[i], h, final_depth(h, depth, which, n, i));
        err_out[i] = (err < tol) ? err : 0.0;
    }
}
// ... and their sum in ascending index order, by one thread
__global__ void __launch_bounds__(kThreads) k_sequential_sum(const double *__restrict__ x, int n, double *out)
{
    __shared__ double buf[2][2048];
    double s = 0.0;
    for (int t = threadIdx.x; t < 2048 && t < n; t += kThreads) buf[0][t] = x[t];
    __syncthreads();
    for (int base = 0, b = 0; base < n; base += 2048, b ^= 1) {
        const int cnt = n - base < 2048 ? n - base : 2048;
        if (threadIdx.x == 0) {
            for (int i = 0; i < cnt; ++i) s += buf[b][i];                 // index order: the reference's loop
        } else {
            for (int t = threadIdx.x - 1; t < 2048 && base + 2048 + t < n; t += kThreads - 1) buf[b ^ 1][t] = x[base + 2048 + t];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = s;
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

// Scores one batch of Hb <= kHChunk hypotheses (host array hyps7) on device-resident points.  The depth-only LM of
// every hypothesis runs without host round trips: [pass, controller] pairs are queued blind -- a finished hypothesis
// makes its CTAs return at once -- followed by the scoring pass; ONE synchronisation then tells whether every solve
// had terminated (almost always: 3-4 passes) or more passes are needed.
static int score_batch(rsdsfm_ctx *ctx, const double *q, const double *u, const double *alpha, const double *alpha_k, int n,
                       const double *hyps7, int Hb, double tol, int *counts, double *sumerr, HypDev **hd_out)
{
    rsdsfm_lm_options opt;
    rsdsfm_lm_default_options(&opt);
    static_assert(kHG == 1, "one hypothesis per CTA row");
    const int gy = Hb;
    const int gx = grid_for(ctx, n, 2);
    constexpr int widthP = RS_NS + RM_NM, widthS = 2;
    RS_TRY(ensure(ctx, ctx->partials, sizeof(double) * (size_t)gx * gy * widthP));
    RS_TRY(ensure(ctx, ctx->sums, sizeof(double) * ((size_t)gy * widthS + 8)));
    RS_TRY(ensure(ctx, ctx->hyp, sizeof(HypDev) * (size_t)kHChunk + sizeof(HypCtl) * (size_t)kHChunk + sizeof(double) * 7 * (size_t)kHChunk + 64));
    RS_TRY(ensure(ctx, ctx->rdepth, sizeof(double) * 2 * (size_t)gy * (size_t)(n > 0 ? n : 1)));
    RS_TRY(ensure_pinned(ctx, sizeof(double) * ((size_t)gy * widthS + 8) + sizeof(double) * 7 * (size_t)kHChunk + 64));
    double *hs = (double *)ctx->pinned;                                   // [gy * 2] counts / error sums, then the active count
    double *h7 = hs + (size_t)gy * widthS + 8;                            // staging of the hypotheses
    HypDev *hd = (HypDev *)ctx->hyp.p;
    HypCtl *hc = (HypCtl *)(hd + kHChunk);
    double *d7 = (double *)(hc + kHChunk);
    int *n_active = (int *)(d7 + 7 * kHChunk);
    double *partials = (double *)ctx->partials.p, *sums = (double *)ctx->sums.p, *depth = (double *)ctx->rdepth.p;
    int *sums_active = (int *)(sums + (size_t)gy * widthS);
    *hd_out = hd;

    memcpy(h7, hyps7, sizeof(double) * 7 * (size_t)Hb);
    RS_CUDA(ctx, cudaMemcpyAsync(d7, h7, sizeof(double) * 7 * (size_t)Hb, cudaMemcpyHostToDevice, ctx->stream));
    RS_CUDA(ctx, cudaMemsetAsync(n_active, 0, sizeof(int), ctx->stream));
    k_ransac_init<<<1, kHChunk, 0, ctx->stream>>>(d7, Hb, n, opt, hd, hc, n_active);
    ctx->launches++;
    const dim3 grid(gy, gx);          // x = hypothesis row (fastest), y = point chunk
    int passes = 4;                   // a depth-only solve takes 3-4 passes
    for (int round = 0; round < 64; ++round) {
        if (n > 0) {
            for (int p = 0; p < passes; ++p) {
                k_ransac_pass<kHG, kPassOcc><<<grid, kThreads, 0, ctx->stream>>>((const double2 *)q, (const double2 *)u, alpha, alpha_k, n, hd,
                                                                                  Hb, opt.min_lm_diagonal, opt.max_lm_diagonal, depth, partials);
                k_ransac_ctl<<<gy, kThreads, 0, ctx->stream>>>(partials, gx, hd, hc, n_active);
            }
            k_ransac_score<kHG, kScoreOcc><<<grid, kThreads, 0, ctx->stream>>>((const double2 *)q, (const double2 *)u, alpha, alpha_k, n, hd, Hb,
                                                                                tol, depth, partials);
            k_ransac_reduce<<<gy, kThreads, 0, ctx->stream>>>(partials, gx, widthS, widthS, sums);
            ctx->launches += 2 * passes + 2;
        } else {
            RS_CUDA(ctx, cudaMemsetAsync(sums, 0, sizeof(double) * (size_t)gy * widthS, ctx->stream));
        }
        RS_CUDA(ctx, cudaMemcpyAsync(sums_active, n_active, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
        RS_CUDA(ctx, cudaMemcpyAsync(hs, sums, sizeof(double) * ((size_t)gy * widthS + 1), cudaMemcpyDeviceToHost, ctx->stream));
        RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (*(const int *)(hs + (size_t)gy * widthS) <= 0) break;        // every solve had terminated before the scoring pass
        passes = 2;
    }
    for (int h = 0; h < Hb; ++h) {
        counts[h] = (int)hs[(size_t)h * widthS];
        sumerr[h] = hs[(size_t)h * widthS + 1];
    }
    return RSDSFM_OK;
}

// minimal.cc:278 decides a tie in the inlier count by `<` on the error sums, which the reference accumulates in index
// order; the parallel sums agree with those to rounding only.  For a tie the sums of both contenders are therefore
// recomputed in that order (rare: per-point errors in parallel, then one thread adds them up).
static int exact_error_sum(rsdsfm_ctx *ctx, const double *q, const double *u, const double *alpha, const double *alpha_k, int n,
                           const HypDev *hd, int which, double tol, double *out)
{
    RS_TRY(ensure(ctx, ctx->misc, sizeof(double) * ((size_t)n + 8)));
    double *err = (double *)ctx->misc.p;
    k_ransac_errors<<<grid_for(ctx, n, 4), kThreads, 0, ctx->stream>>>((const double2 *)q, (const double2 *)u, alpha, alpha_k, n, hd, which, tol,
                                                                       (const double *)ctx->rdepth.p, err);
    k_sequential_sum<<<1, kThreads, 0, ctx->stream>>>(err, n, err + n);
    ctx->launches += 2;
    RS_TRY(ensure_pinned(ctx, 64));
    RS_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, err + n, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = *(const double *)ctx->pinned;
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
        // minimal.cc:278 scans the hypotheses in order and replaces the leader on a larger count, or on an equal count
        // and a smaller error sum: the winner is, among the hypotheses with the largest count, the first one with the
        // smallest sum.  Only ties AT the largest count need the index-order sums.
        int batch_max = -1, n_max = 0;
        for (int h = 0; h < Hb; ++h) {
            if (counts) counts[b0 + h] = cnt[(size_t)h];
            if (sumerr) sumerr[b0 + h] = err[(size_t)h];
            if (cnt[(size_t)h] > batch_max) { batch_max = cnt[(size_t)h]; n_max = 1; }
            else if (cnt[(size_t)h] == batch_max) ++n_max;
        }
        // The parallel sums agree with the index-order sums to n * 2^-53 relative at worst (n < 2^23: 1e-9), so only
        // contenders whose parallel sums lie within that of the smallest one can change places: those are re-summed
        // in index order (rare: duplicated or scaled copies of one hypothesis), everybody else is decided already.
        double min_par = 0.0;
        bool have = false;
        for (int h = 0; h < Hb; ++h)
            if (cnt[(size_t)h] == batch_max && (!have || err[(size_t)h] < min_par)) { min_par = err[(size_t)h]; have = true; }
        const double band = min_par + 1e-9 * fabs(min_par);
        int n_close = 0;
        for (int h = 0; h < Hb; ++h) if (cnt[(size_t)h] == batch_max && err[(size_t)h] <= band) ++n_close;
        int batch_best = -1;
        double batch_err = 0.0;
        for (int h = 0; h < Hb; ++h) {
            if (cnt[(size_t)h] != batch_max || !(err[(size_t)h] <= band)) continue;
            if (n_close > 1 && batch_max > 0 && n > 0) {
                RS_TRY(exact_error_sum(ctx, q, u, alpha, alpha_k, n, hd, h, tol, &err[(size_t)h]));
                if (sumerr) sumerr[b0 + h] = err[(size_t)h];             // report the index-order sums where they were formed
            }
            if (batch_best < 0 || err[(size_t)h] < batch_err) { batch_best = h; batch_err = err[(size_t)h]; }
        }
        // against the leader of the earlier batches (its depths are gone: a tie across batches is decided on the sums as they are)
        if (batch_best >= 0 && (batch_max > best_count || (batch_max == best_count && batch_err < best_err))) {
            best_count = batch_max; best_err = batch_err; best = b0 + batch_best;
        } else {
            batch_best = -1;
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
