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
// decisions, which are functions of a few global sums.  So nothing per (point, hypothesis) is
// stored: a pass REPLAYS the accepted steps of a hypothesis from d = 1 (radii kept per
// hypothesis), evaluates the next candidate and reduces the sums the controller needs.
// Algorithmic traffic: 48 B per point per pass per group of kHG hypotheses.
#include "common.cuh"
#include "lm_controller.h"
#include "solve9.h"

namespace rsdsfm {

constexpr int kHG = 4;          // hypotheses evaluated per point load
constexpr int kMaxAcc = 50;     // >= max_num_iterations

struct HypDev {
    double w[3], v[3], k;
    double radius;              // radius of the candidate step of this pass
    int n_acc;                  // accepted steps so far
    int active;                 // still iterating
    int failed;                 // solver FAILURE: depths keep their start value 1.0
    int pad;
    double acc_radius[kMaxAcc];
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
    const double beta = (2.0 / (2.0 + h.k)) * (alpha + h.k * alpha_k);
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

__device__ __forceinline__ double depth_replay(const DepthTerms &T, const HypDev &h)
{
    double d = 1.0;                                         // nonlinearRefinement.cc:140
    if (h.failed) return d;
    for (int a = 0; a < h.n_acc; ++a) {
        double r0, r1;
        depth_residual(T, d, r0, r1);
        d = d + depth_step(T, r0, r1, h.acc_radius[a]) * T.scale;
    }
    return d;
}

__global__ void __launch_bounds__(kThreads) k_ransac_pass(const double2 *__restrict__ q, const double2 *__restrict__ u,
                                                          const double *__restrict__ alpha,
                                                          const double *__restrict__ alpha_k, int n,
                                                          const HypDev *__restrict__ hyps, int H, double min_diag,
                                                          double max_diag, double *__restrict__ partials)
{
    __shared__ HypDev sh[kHG];
    const int h0 = blockIdx.y * kHG;
    for (int t = threadIdx.x; t < (int)(kHG * sizeof(HypDev) / sizeof(int)); t += blockDim.x) {
        const int hh = t / (int)(sizeof(HypDev) / sizeof(int));
        int val = 0;
        if (h0 + hh < H) val = reinterpret_cast<const int *>(hyps + h0)[t];
        reinterpret_cast<int *>(sh)[t] = val;
    }
    __syncthreads();
    double s[kHG * RS_NS], mx[kHG * RM_NM];
#pragma unroll
    for (int j = 0; j < kHG * RS_NS; ++j) s[j] = 0.0;
#pragma unroll
    for (int j = 0; j < kHG * RM_NM; ++j) mx[j] = 0.0;

    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double2 qq = q[i], uu = u[i];
        const double al = alpha[i], alk = alpha_k[i];
#pragma unroll
        for (int g = 0; g < kHG; ++g) {
            const HypDev &h = sh[g];
            if (!h.active) continue;
            DepthTerms T;
            depth_terms(qq.x, qq.y, uu.x, uu.y, al, alk, h, min_diag, max_diag, T);
            const double d = depth_replay(T, h);
            double r0, r1;
            depth_residual(T, d, r0, r1);
            // evaluation at x
            s[g * RS_NS + RS_COST] += 0.5 * (r0 * r0 + r1 * r1);
            s[g * RS_NS + RS_SUMSQ_D] += d * d;
            const double ge = T.E0 * r0 + T.E1 * r1;
            const double proj = d + (-ge);
            mx[g * RM_NM + RM_GMAX_E] = fmax(mx[g * RM_NM + RM_GMAX_E], fabs(d - proj));
            mx[g * RM_NM + RM_BAD] = fmax(mx[g * RM_NM + RM_BAD], bad_flag(r0) + bad_flag(r1) + bad_flag(T.E0) + bad_flag(T.E1));
            // candidate step
            const double step_e = depth_step(T, r0, r1, h.radius);
            const double mr0 = T.e0 * step_e, mr1 = T.e1 * step_e;
            s[g * RS_NS + RS_MCC] += mr0 * (r0 + mr0 / 2.0) + mr1 * (r1 + mr1 / 2.0);
            const double dc = d + step_e * T.scale;
            const double dd = d - dc;
            s[g * RS_NS + RS_STEP_SQ] += dd * dd;
            double c0, c1;
            depth_residual(T, dc, c0, c1);
            s[g * RS_NS + RS_CAND_COST] += 0.5 * (c0 * c0 + c1 * c1);
            mx[g * RM_NM + RM_BAD_STEP] = fmax(mx[g * RM_NM + RM_BAD_STEP], bad_flag(step_e));
            mx[g * RM_NM + RM_BAD_CAND] = fmax(mx[g * RM_NM + RM_BAD_CAND], bad_flag(c0) + bad_flag(c1));
        }
    }
    block_reduce_store<kHG * RS_NS, kHG * RM_NM>(s, mx, partials + (size_t)blockIdx.y * gridDim.x * (kHG * (RS_NS + RM_NM)));
}

// Scoring loop minimal.cc:255-275 at the final depths: per hypothesis inlier count and error sum.
__device__ __forceinline__ double ransac_error(double x, double y, double ux, double uy, double alpha, double alpha_k,
                                               const HypDev &h, double d)
{
    const double av0 = 1.0 * h.v[0] + 0.0 * h.v[1] + (-x) * h.v[2];
    const double av1 = 0.0 * h.v[0] + 1.0 * h.v[1] + (-y) * h.v[2];
    const double bw0 = (-x * y) * h.w[0] + (1 + x * x) * h.w[1] + (-y) * h.w[2];
    const double bw1 = (-(1 + y * y)) * h.w[0] + (x * y) * h.w[1] + x * h.w[2];
    const double beta = (alpha + h.k * alpha_k) * (2.0 / (2.0 + h.k));
    const double ue0 = beta * (av0 * d + bw0);
    const double ue1 = beta * (av1 * d + bw1);
    const double dx = ue0 - ux, dy = ue1 - uy;
    return sqrt(dx * dx + dy * dy);
}

__global__ void __launch_bounds__(kThreads) k_ransac_score(const double2 *__restrict__ q, const double2 *__restrict__ u,
                                                           const double *__restrict__ alpha,
                                                           const double *__restrict__ alpha_k, int n,
                                                           const HypDev *__restrict__ hyps, int H, double min_diag,
                                                           double max_diag, double tol, double *__restrict__ partials)
{
    __shared__ HypDev sh[kHG];
    const int h0 = blockIdx.y * kHG;
    for (int t = threadIdx.x; t < (int)(kHG * sizeof(HypDev) / sizeof(int)); t += blockDim.x) {
        const int hh = t / (int)(sizeof(HypDev) / sizeof(int));
        int val = 0;
        if (h0 + hh < H) val = reinterpret_cast<const int *>(hyps + h0)[t];
        reinterpret_cast<int *>(sh)[t] = val;
    }
    __syncthreads();
    double s[kHG * 2], mx[1] = {0.0};
#pragma unroll
    for (int j = 0; j < kHG * 2; ++j) s[j] = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double2 qq = q[i], uu = u[i];
        const double al = alpha[i], alk = alpha_k[i];
#pragma unroll
        for (int g = 0; g < kHG; ++g) {
            if (h0 + g >= H) continue;
            const HypDev &h = sh[g];
            DepthTerms T;
            depth_terms(qq.x, qq.y, uu.x, uu.y, al, alk, h, min_diag, max_diag, T);
            const double d = depth_replay(T, h);
            const double err = ransac_error(qq.x, qq.y, uu.x, uu.y, al, alk, h, d);
            if (err < tol) { s[g * 2] += 1.0; s[g * 2 + 1] += err; }
        }
    }
    block_reduce_store<kHG * 2, 0>(s, mx, partials + (size_t)blockIdx.y * gridDim.x * (kHG * 2));
}

__global__ void __launch_bounds__(kThreads) k_ransac_winner(const double2 *__restrict__ q, const double2 *__restrict__ u,
                                                            const double *__restrict__ alpha,
                                                            const double *__restrict__ alpha_k, int n,
                                                            const HypDev *__restrict__ hyps, int best, double min_diag,
                                                            double max_diag, double tol, uint8_t *mask, double *inv_depth)
{
    __shared__ HypDev h;
    for (int t = threadIdx.x; t < (int)(sizeof(HypDev) / sizeof(int)); t += blockDim.x)
        reinterpret_cast<int *>(&h)[t] = reinterpret_cast<const int *>(hyps + best)[t];
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double2 qq = q[i], uu = u[i];
        const double al = alpha[i], alk = alpha_k[i];
        DepthTerms T;
        depth_terms(qq.x, qq.y, uu.x, uu.y, al, alk, h, min_diag, max_diag, T);
        const double d = depth_replay(T, h);
        const double err = ransac_error(qq.x, qq.y, uu.x, uu.y, al, alk, h, d);
        inv_depth[i] = d;
        mask[i] = (err < tol) ? 1 : 0;
    }
}

// rows of `width` doubles per CTA, grid (gx, gy): out[by*width + j] = sum/max over bx ascending
__global__ void k_ransac_reduce(const double *__restrict__ partials, int gx, int width, int nsum, double *out)
{
    const int by = blockIdx.x;
    for (int j = threadIdx.x; j < width; j += blockDim.x) {
        const double *p = partials + (size_t)by * gx * width + j;
        double v = p[0];
        if (j < nsum) for (int b = 1; b < gx; ++b) v += p[(size_t)b * width];
        else          for (int b = 1; b < gx; ++b) v = fmax(v, p[(size_t)b * width]);
        out[(size_t)by * width + j] = v;
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

// Scores H hypotheses (host array hyps7) on device-resident points.
int ransac_score_device(rsdsfm_ctx *ctx, const double *q, const double *u, const double *alpha, const double *alpha_k,
                        int n, const double *hyps7, int H, double tol, int *counts, double *sumerr, int *best_idx,
                        uint8_t *mask_best, double *inv_depth_best)
{
    if (H <= 0) { *best_idx = -1; return RSDSFM_OK; }
    rsdsfm_lm_options opt;
    rsdsfm_lm_default_options(&opt);
    const int gy = (H + kHG - 1) / kHG;
    int gx = grid_for(ctx, n, 2);
    const int widthP = kHG * (RS_NS + RM_NM), widthS = kHG * 2;
    RS_TRY(ensure(ctx, ctx->partials, sizeof(double) * (size_t)gx * gy * widthP));
    RS_TRY(ensure(ctx, ctx->sums, sizeof(double) * (size_t)gy * widthP));
    RS_TRY(ensure(ctx, ctx->hyp, sizeof(HypDev) * (size_t)gy * kHG));
    RS_TRY(ensure_pinned(ctx, sizeof(HypDev) * (size_t)gy * kHG + sizeof(double) * (size_t)gy * widthP + 64));
    HypDev *hh = (HypDev *)ctx->pinned;
    double *hs = (double *)((char *)ctx->pinned + sizeof(HypDev) * (size_t)gy * kHG);
    HypDev *hd = (HypDev *)ctx->hyp.p;
    double *partials = (double *)ctx->partials.p, *sums = (double *)ctx->sums.p;

    std::vector<LmController> ctl((size_t)H);
    std::vector<LmNext> pend((size_t)H, LM_RUN_A);
    memset(hh, 0, sizeof(HypDev) * (size_t)gy * kHG);
    int n_active = 0;
    for (int h = 0; h < H; ++h) {
        const double *p = hyps7 + 7 * h;
        for (int j = 0; j < 3; ++j) { hh[h].w[j] = p[j]; hh[h].v[j] = p[3 + j]; }
        hh[h].k = p[6];
        ctl[h].init(opt, 0, nullptr);
        hh[h].radius = ctl[h].radius;
        bool finite = true;
        for (int j = 0; j < 7; ++j) finite = finite && isfinite(p[j]);
        // solver.cc: non-finite parameter values => FAILURE, depths stay at 1.0.  n == 0: nothing to do.
        hh[h].active = (finite && n > 0) ? 1 : 0;
        hh[h].failed = finite ? 0 : 1;
        n_active += hh[h].active;
    }
    const dim3 grid(gx, gy);
    while (n_active > 0) {
        RS_CUDA(ctx, cudaMemcpyAsync(hd, hh, sizeof(HypDev) * (size_t)gy * kHG, cudaMemcpyHostToDevice, ctx->stream));
        k_ransac_pass<<<grid, kThreads, 0, ctx->stream>>>((const double2 *)q, (const double2 *)u, alpha, alpha_k, n, hd, H,
                                                          opt.min_lm_diagonal, opt.max_lm_diagonal, partials);
        k_ransac_reduce<<<gy, 64, 0, ctx->stream>>>(partials, gx, widthP, kHG * RS_NS, sums);
        ctx->launches += 2;
        RS_CUDA(ctx, cudaMemcpyAsync(hs, sums, sizeof(double) * (size_t)gy * widthP, cudaMemcpyDeviceToHost, ctx->stream));
        RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        n_active = 0;
        for (int h = 0; h < H; ++h) {
            if (!hh[h].active) continue;
            const double *row = hs + (size_t)(h / kHG) * widthP;
            const double *ss = row + (h % kHG) * RS_NS, *mm = row + kHG * RS_NS + (h % kHG) * RM_NM;
            EvalSums e;
            memset(&e, 0, sizeof e);
            e.cost = ss[RS_COST]; e.sumsq_d = ss[RS_SUMSQ_D]; e.gmax_e = mm[RM_GMAX_E]; e.bad = mm[RM_BAD];
            CandSums c;
            c.mcc = ss[RS_MCC]; c.step_sq = ss[RS_STEP_SQ]; c.cand_cost = ss[RS_CAND_COST];
            c.bad_step = mm[RM_BAD_STEP]; c.bad_cand = mm[RM_BAD_CAND];
            const double used_radius = hh[h].radius;
            LmNext next = pend[h];
            if (next == LM_RUN_A) next = ctl[h].on_eval(e);               // evaluation at a new point
            while (next == LM_SOLVE) next = ctl[h].solve_step(nullptr);   // no f-blocks: nothing to solve
            if (next == LM_RUN_B) next = ctl[h].on_candidate(c);
            pend[h] = next;
            if (next == LM_DONE) {
                hh[h].active = 0;
                hh[h].failed = (ctl[h].termination == RSDSFM_FAILURE) ? 1 : 0;
            } else {
                if (ctl[h].accepted_last && hh[h].n_acc < kMaxAcc) hh[h].acc_radius[hh[h].n_acc++] = used_radius;
                hh[h].radius = ctl[h].radius;
                n_active++;
            }
        }
    }
    // scoring
    RS_CUDA(ctx, cudaMemcpyAsync(hd, hh, sizeof(HypDev) * (size_t)gy * kHG, cudaMemcpyHostToDevice, ctx->stream));
    if (n > 0) {
        k_ransac_score<<<grid, kThreads, 0, ctx->stream>>>((const double2 *)q, (const double2 *)u, alpha, alpha_k, n, hd, H,
                                                           opt.min_lm_diagonal, opt.max_lm_diagonal, tol, partials);
        k_ransac_reduce<<<gy, 64, 0, ctx->stream>>>(partials, gx, widthS, widthS, sums);
        ctx->launches += 2;
        RS_CUDA(ctx, cudaMemcpyAsync(hs, sums, sizeof(double) * (size_t)gy * widthS, cudaMemcpyDeviceToHost, ctx->stream));
        RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int j = 0; j < gy * widthS; ++j) hs[j] = 0.0;
    }
    int best = -1, best_count = -1;
    double best_err = 0.0;
    for (int h = 0; h < H; ++h) {
        const double *row = hs + (size_t)(h / kHG) * widthS + (h % kHG) * 2;
        const int c = (int)row[0];
        const double e = row[1];
        if (counts) counts[h] = c;
        if (sumerr) sumerr[h] = e;
        if (c > best_count || (c == best_count && e < best_err)) { best_count = c; best_err = e; best = h; }   // :278
    }
    *best_idx = best;
    if (best >= 0 && n > 0 && (mask_best || inv_depth_best)) {
        k_ransac_winner<<<grid_for(ctx, n, 4), kThreads, 0, ctx->stream>>>((const double2 *)q, (const double2 *)u, alpha,
                                                                           alpha_k, n, hd, best, opt.min_lm_diagonal,
                                                                           opt.max_lm_diagonal, tol, mask_best, inv_depth_best);
        ctx->launches++;
    }
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
