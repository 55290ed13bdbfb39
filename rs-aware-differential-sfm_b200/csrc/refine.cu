// refine.cu -- GPU Levenberg-Marquardt for the reference's two Ceres problems:
//   a8  nonlinear_refinement::estimateInverseDepths  (nonlinearRefinement.cc:109-180)  NF = 0
//   a9  nonlinear_refinement::nonLinearRefinement    (nonlinearRefinement.cc:183-252)  NF = 6 | 7
//
// Every per-pixel inverse depth is a 1x1 Schur e-block that is eliminated in closed form inside
// the pass that evaluates the residual and Jacobian (pass A); the small dense motion system is
// accumulated in FP64 (registers -> warp shuffles -> shared memory -> one row per CTA -> fixed
// order final reduce).  Pass B back-substitutes the depths, forms the candidate point and its
// cost, the model cost change and the step norm.  lm_controller.h holds the O(1) trust-region
// logic (Ceres 1.14 semantics).
//
// Data layout in HBM (structure of arrays, one entry per residual block, coalesced):
//   xy[m]  double2 (x, y)        normalised coordinates of inlier i
//   uu[m]  double2 (ux, uy)      gamma-scaled normalised flow paired with it (Q1 pairing applied
//                                 once, in the gather kernel)
//   aa[m]  double2 (alpha, alpha_k)
//   d[2][m] double               inverse depth, ping-pong (x / candidate)
//   se[m]  double                Jacobi scale of the depth column, fixed at iteration 0
#include "common.cuh"
#include "lm_controller.h"
#include "rs_math.cuh"

namespace rsdsfm {

struct RefineData {
    const double2 *xy;
    const double2 *uu;
    const double2 *aa;
    double *se;
    int m;
};

// ------------------------------------------------------------------------------------------
// gather: API arrays -> SoA records.  Residual i pairs inlier i with flow(:, i) of the array the
// caller passed (reference behaviour, nonlinearRefinement.cc:209-212) or flow(:, flow_index[i]).
// ------------------------------------------------------------------------------------------
__global__ void k_refine_gather(const double *__restrict__ flow, const double *__restrict__ inliers3,
                                const double *__restrict__ alpha, const double *__restrict__ alpha_k,
                                const int32_t *__restrict__ flow_index, int m, double2 *xy, double2 *uu, double2 *aa,
                                double *d0)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const double x = inliers3[3 * (size_t)i], y = inliers3[3 * (size_t)i + 1], z = inliers3[3 * (size_t)i + 2];
        const int fi = flow_index ? flow_index[i] : i;
        const double2 u = reinterpret_cast<const double2 *>(flow)[fi];
        xy[i] = make_double2(x, y);
        uu[i] = u;
        aa[i] = make_double2(alpha[i], alpha_k[i]);
        d0[i] = 1.0 / z;                                   // nonlinearRefinement.cc:213
    }
}

// a8 variant: coordinates / flow already interleaved pairs, depth starts at 1.0 (:140)
__global__ void k_depth_gather(const double *__restrict__ alpha, const double *__restrict__ alpha_k, int n,
                               double2 *aa, double *d0)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        aa[i] = make_double2(alpha[i], alpha_k[i]);
        d0[i] = 1.0;
    }
}

__global__ void k_check_finite(const double *__restrict__ d, int m, int *flag)
{
    int bad = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
        if (!isfinite(d[i])) bad = 1;
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flag, 1);
}

__device__ __forceinline__ Obs load_obs(const RefineData &D, int i)
{
    const double2 p = D.xy[i], u = D.uu[i], a = D.aa[i];
    Obs o;
    o.x = p.x; o.y = p.y; o.ux = u.x; o.uy = u.y; o.alpha = a.x; o.alpha_k = a.y;
    return o;
}

struct PassParams {
    Motion mot;        // current x (motion part)
    Motion cand;       // candidate motion (pass B)
    double delta_f[kMaxNF];
    double radius;
    double min_diag, max_diag;
    int first;         // pass A of iteration 0: compute and store the depth-column Jacobi scale
};

// ------------------------------------------------------------------------------------------
// Pass A: residual + Jacobian at x, 1x1 Schur elimination of each depth, FP64 accumulation of
// the reduced system.  One row of SumsA::NS + SumsA::NM doubles per CTA.
// ------------------------------------------------------------------------------------------
template <int NF>
__global__ void __launch_bounds__(kThreads) k_lm_pass_a(RefineData D, const double *__restrict__ d, PassParams P,
                                                        double *__restrict__ partials)
{
    double s[SumsA::NS];
    double mx[SumsA::NM];
#pragma unroll
    for (int j = 0; j < SumsA::NS; ++j) s[j] = 0.0;
    mx[0] = 0.0; mx[1] = 0.0;
    const double c2 = 2.0 / (2.0 + P.mot.k);
    const double inv_radius = 1.0 / P.radius;

    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D.m; i += gridDim.x * blockDim.x) {
        const Obs o = load_obs(D, i);
        const double di = d[i];
        double r0, r1, e0, e1, F0[NF > 0 ? NF : 1], F1[NF > 0 ? NF : 1];
        rs_residual_jac<NF>(o, P.mot, c2, di, r0, r1, e0, e1, F0, F1);
        const double ee = e0 * e0 + e1 * e1;
        double se;
        if (P.first) { se = 1.0 / (1.0 + sqrt(ee)); D.se[i] = se; }   // jacobian_scaling_, fixed at iteration 0
        else se = D.se[i];
        double bad = bad_flag(r0) + bad_flag(r1) + bad_flag(ee);
        s[SumsA::COST] += 0.5 * (r0 * r0 + r1 * r1);
        s[SumsA::SUMSQ_D] += di * di;
        const double ge = e0 * r0 + e1 * r1;                         // gradient of the depth block
        mx[SumsA::GMAX_E] = fmax(mx[SumsA::GMAX_E], fabs(di - (di - ge)));
        if (NF > 0) {
            // e-block: ete = e_s^T e_s + D_e^2, D_e^2 = clamp(e_s^T e_s)/radius; q = s_e^2 / ete
            const double ees = ee * se * se;
            const double ete = ees + fmin(fmax(ees, P.min_diag), P.max_diag) * inv_radius;
            const double q = se * se / ete;
            double fe[NF > 0 ? NF : 1];                              // F^T e
#pragma unroll
            for (int j = 0; j < NF; ++j) {
                fe[j] = F0[j] * e0 + F1[j] * e1;
                bad += bad_flag(fe[j]);
                s[SumsA::GF + j] += F0[j] * r0 + F1[j] * r1;
                s[SumsA::CSF + j] += F0[j] * F0[j] + F1[j] * F1[j];
            }
            const double qge = q * ge;
            int t = 0;
#pragma unroll
            for (int j = 0; j < NF; ++j) {
                s[SumsA::RHS + j] += F0[j] * r0 + F1[j] * r1 - fe[j] * qge;
                const double qf = q * fe[j];
#pragma unroll
                for (int c = j; c < NF; ++c, ++t)
                    s[SumsA::S + t] += F0[j] * F0[c] + F1[j] * F1[c] - qf * fe[c];
            }
        }
        mx[SumsA::BAD] = fmax(mx[SumsA::BAD], bad);
    }
    block_reduce_store<SumsA::NS, SumsA::NM>(s, mx, partials);
}

// ------------------------------------------------------------------------------------------
// Pass B: back substitution of every depth, candidate point, candidate cost, model cost change,
// squared step norm.
// ------------------------------------------------------------------------------------------
template <int NF>
__global__ void __launch_bounds__(kThreads) k_lm_pass_b(RefineData D, const double *__restrict__ d,
                                                        double *__restrict__ d_cand, PassParams P,
                                                        double *__restrict__ partials)
{
    double s[SumsB::NS] = {0.0, 0.0, 0.0};
    double mx[SumsB::NM] = {0.0, 0.0};
    const double c2 = 2.0 / (2.0 + P.mot.k);
    const double c2c = 2.0 / (2.0 + P.cand.k);
    const double inv_radius = 1.0 / P.radius;

    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D.m; i += gridDim.x * blockDim.x) {
        const Obs o = load_obs(D, i);
        const double di = d[i];
        const double se = D.se[i];
        double r0, r1, e0, e1, F0[NF > 0 ? NF : 1], F1[NF > 0 ? NF : 1];
        rs_residual_jac<NF>(o, P.mot, c2, di, r0, r1, e0, e1, F0, F1);
        const double ees = (e0 * e0 + e1 * e1) * se * se;
        const double ete = ees + fmin(fmax(ees, P.min_diag), P.max_diag) * inv_radius;
        const double q = se * se / ete;
        // F delta_f
        double m0 = 0.0, m1 = 0.0;
#pragma unroll
        for (int j = 0; j < NF; ++j) { m0 += F0[j] * P.delta_f[j]; m1 += F1[j] * P.delta_f[j]; }
        // delta_e = -q e^T (r + F delta_f)
        const double delta_e = -q * (e0 * (r0 + m0) + e1 * (r1 + m1));
        m0 += e0 * delta_e; m1 += e1 * delta_e;                       // J delta
        s[SumsB::MCC] += m0 * (r0 + 0.5 * m0) + m1 * (r1 + 0.5 * m1);
        const double dc = di + delta_e;
        const double dd = di - dc;
        s[SumsB::STEP_SQ] += dd * dd;
        d_cand[i] = dc;
        double c0, c1;
        rs_residual(o, P.cand, c2c, dc, c0, c1);
        s[SumsB::CAND_COST] += 0.5 * (c0 * c0 + c1 * c1);
        mx[SumsB::BAD_STEP] = fmax(mx[SumsB::BAD_STEP], bad_flag(delta_e));
        mx[SumsB::BAD_CAND] = fmax(mx[SumsB::BAD_CAND], bad_flag(c0) + bad_flag(c1));
    }
    block_reduce_store<SumsB::NS, SumsB::NM>(s, mx, partials);
}

__global__ void k_final_reduce(const double *__restrict__ partials, int nblocks, int ns, int nm, double *out)
{
    const int j = threadIdx.x;
    if (j >= ns + nm) return;
    double v = partials[j];
    if (j < ns) for (int b = 1; b < nblocks; ++b) v += partials[(size_t)b * (ns + nm) + j];
    else        for (int b = 1; b < nblocks; ++b) v = fmax(v, partials[(size_t)b * (ns + nm) + j]);
    out[j] = v;
}

void launch_final_reduce(rsdsfm_ctx *ctx, const double *partials, int nblocks, int ns, int nm, double *out)
{
    k_final_reduce<<<1, 64, 0, ctx->stream>>>(partials, nblocks, ns, nm, out);
    ctx->launches++;
}

__global__ void k_invert(const double *__restrict__ d, int m, double *__restrict__ z)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) z[i] = 1.0 / d[i];
}
__global__ void k_copy(const double *__restrict__ d, int m, double *__restrict__ z)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) z[i] = d[i];
}

static void motion_from_f(int nf, const double *f, const Motion &base, Motion &out)
{
    out = base;
    if (nf >= 6) { for (int j = 0; j < 3; ++j) { out.v[j] = f[j]; out.w[j] = f[3 + j]; } }
    if (nf == 7) out.k = f[6];
}

template <int NF>
static void launch_a(rsdsfm_ctx *ctx, int grid, const RefineData &D, const double *d, const PassParams &P, double *partials)
{
    k_lm_pass_a<NF><<<grid, kThreads, 0, ctx->stream>>>(D, d, P, partials);
    ctx->launches++;
}
template <int NF>
static void launch_b(rsdsfm_ctx *ctx, int grid, const RefineData &D, const double *d, double *dc, const PassParams &P, double *partials)
{
    k_lm_pass_b<NF><<<grid, kThreads, 0, ctx->stream>>>(D, d, dc, P, partials);
    ctx->launches++;
}

// The LM solve on device-resident SoA data.  d[0] holds the start depths; on return *d_final
// points at the buffer holding the result.  base = fixed motion values; nf selects free blocks.
int lm_solve_device(rsdsfm_ctx *ctx, const RefineData &D, double *dbuf0, double *dbuf1, int nf, Motion &mot,
                    const rsdsfm_lm_options &opt, rsdsfm_lm_summary *summary, double **d_final)
{
    const int grid = grid_for(ctx, D.m);
    const int rowA = SumsA::NS + SumsA::NM, rowB = SumsB::NS + SumsB::NM;
    RS_TRY(ensure(ctx, ctx->partials, sizeof(double) * (size_t)grid * rowA));
    RS_TRY(ensure(ctx, ctx->sums, sizeof(double) * rowA));
    RS_TRY(ensure_pinned(ctx, sizeof(double) * rowA + 64));
    double *partials = (double *)ctx->partials.p, *sums = (double *)ctx->sums.p;
    double *hs = (double *)ctx->pinned;

    double f0[kMaxNF] = {0, 0, 0, 0, 0, 0, 0};
    if (nf >= 6) { for (int j = 0; j < 3; ++j) { f0[j] = mot.v[j]; f0[3 + j] = mot.w[j]; } }
    if (nf == 7) f0[6] = mot.k;
    LmController ctl;
    ctl.init(opt, nf, f0);
    const Motion base = mot;
    double *dx = dbuf0, *dc = dbuf1;

    PassParams P;
    P.min_diag = opt.min_lm_diagonal; P.max_diag = opt.max_lm_diagonal;
    RS_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    LmNext next = LM_RUN_A;
    bool first = true;
    while (next != LM_DONE) {
        motion_from_f(nf, ctl.f, base, P.mot);
        P.radius = ctl.radius;
        if (next == LM_RUN_A) {
            P.first = first ? 1 : 0;
            if (nf == 0) launch_a<0>(ctx, grid, D, dx, P, partials);
            else if (nf == 6) launch_a<6>(ctx, grid, D, dx, P, partials);
            else launch_a<7>(ctx, grid, D, dx, P, partials);
            first = false;
            launch_final_reduce(ctx, partials, grid, SumsA::NS, SumsA::NM, sums);
            RS_CUDA(ctx, cudaMemcpyAsync(hs, sums, sizeof(double) * rowA, cudaMemcpyDeviceToHost, ctx->stream));
            RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            next = ctl.after_A(hs, hs + SumsA::NS);
        } else {
            for (int j = 0; j < kMaxNF; ++j) P.delta_f[j] = ctl.delta_f[j];
            double fc[kMaxNF];
            for (int j = 0; j < kMaxNF; ++j) fc[j] = ctl.f[j] + ctl.delta_f[j];
            motion_from_f(nf, fc, base, P.cand);
            if (nf == 0) launch_b<0>(ctx, grid, D, dx, dc, P, partials);
            else if (nf == 6) launch_b<6>(ctx, grid, D, dx, dc, P, partials);
            else launch_b<7>(ctx, grid, D, dx, dc, P, partials);
            launch_final_reduce(ctx, partials, grid, SumsB::NS, SumsB::NM, sums);
            RS_CUDA(ctx, cudaMemcpyAsync(hs, sums, sizeof(double) * rowB, cudaMemcpyDeviceToHost, ctx->stream));
            RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            next = ctl.after_B(hs, hs + SumsB::NS);
            if (ctl.accepted_last) { double *t = dx; dx = dc; dc = t; }
        }
    }
    RS_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    RS_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    RS_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (summary) { ctl.fill_summary(summary); summary->device_ms = ms; }
    if (ctl.termination == RSDSFM_FAILURE) {
        mot = base;           // solver.cc Minimize(): original parameters are restored on FAILURE
        *d_final = nullptr;   // caller restores the start depths
    } else {
        motion_from_f(nf, ctl.f, base, mot);
        *d_final = dx;
    }
    return RSDSFM_OK;
}

// a9 on device pointers.  z_out may alias nothing else.
int refine_device(rsdsfm_ctx *ctx, const double *flow, const double *inliers3, const double *alpha,
                  const double *alpha_k, int m, double *v, double *w, double *k, int const_acc,
                  const int32_t *flow_index, const rsdsfm_lm_options *opts, double *z_out, rsdsfm_lm_summary *summary)
{
    rsdsfm_lm_options o;
    if (opts) o = *opts; else rsdsfm_lm_default_options(&o);
    rsdsfm_lm_summary local;
    if (!summary) summary = &local;
    memset(summary, 0, sizeof *summary);
    bool finite = isfinite(*k);
    for (int j = 0; j < 3; ++j) finite = finite && isfinite(v[j]) && isfinite(w[j]);
    if (m == 0) { summary->termination = RSDSFM_CONVERGENCE; summary->reason = RSDSFM_REASON_FUNCTION_TOL; return RSDSFM_OK; }

    const size_t mm = (size_t)m;
    RS_TRY(ensure(ctx, ctx->pix, sizeof(double2) * 3 * mm));
    RS_TRY(ensure(ctx, ctx->dA, sizeof(double) * mm));
    RS_TRY(ensure(ctx, ctx->dB, sizeof(double) * mm));
    RS_TRY(ensure(ctx, ctx->scale_e, sizeof(double) * mm));
    RS_TRY(ensure(ctx, ctx->flags, 64));
    double2 *xy = (double2 *)ctx->pix.p, *uu = xy + mm, *aa = uu + mm;
    double *d0 = (double *)ctx->dA.p, *d1 = (double *)ctx->dB.p;
    const int grid = grid_for(ctx, m, 8);
    k_refine_gather<<<grid, kThreads, 0, ctx->stream>>>(flow, inliers3, alpha, alpha_k, flow_index, m, xy, uu, aa, d0);
    ctx->launches++;
    // solver.cc: non-finite initial parameter values => FAILURE before any evaluation
    RS_CUDA(ctx, cudaMemsetAsync(ctx->flags.p, 0, 4, ctx->stream));
    k_check_finite<<<grid, kThreads, 0, ctx->stream>>>(d0, m, (int *)ctx->flags.p);
    ctx->launches++;
    RS_TRY(ensure_pinned(ctx, 1024));
    RS_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, ctx->flags.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*(int *)ctx->pinned) finite = false;
    double *dfin = nullptr;
    if (!finite) {
        summary->termination = RSDSFM_FAILURE; summary->reason = RSDSFM_REASON_NONFINITE_INPUT;
    } else {
        RefineData D{xy, uu, aa, (double *)ctx->scale_e.p, m};
        Motion mot;
        for (int j = 0; j < 3; ++j) { mot.v[j] = v[j]; mot.w[j] = w[j]; }
        mot.k = *k;
        RS_TRY(lm_solve_device(ctx, D, d0, d1, const_acc ? 7 : 6, mot, o, summary, &dfin));
        for (int j = 0; j < 3; ++j) { v[j] = mot.v[j]; w[j] = mot.w[j]; }
        *k = mot.k;
    }
    if (dfin == nullptr) {
        // FAILURE: the depths keep their start values 1/z (round trip 1/(1/z) like the reference, :213,:247)
        k_refine_gather<<<grid, kThreads, 0, ctx->stream>>>(flow, inliers3, alpha, alpha_k, flow_index, m, xy, uu, aa, d0);
        ctx->launches++;
        dfin = d0;
    }
    k_invert<<<grid, kThreads, 0, ctx->stream>>>(dfin, m, z_out);       // nonlinearRefinement.cc:247
    ctx->launches++;
    return RSDSFM_OK;
}

// a8 on device pointers: coord / flow are interleaved pairs already.
int estimate_inverse_depths_device(rsdsfm_ctx *ctx, const double *coord, const double *flow, int n, const double *v,
                                   const double *w, double k, const double *alpha, const double *alpha_k,
                                   double *inv_depth, rsdsfm_lm_summary *summary)
{
    rsdsfm_lm_options o;
    rsdsfm_lm_default_options(&o);
    rsdsfm_lm_summary local;
    if (!summary) summary = &local;
    memset(summary, 0, sizeof *summary);
    if (n == 0) { summary->termination = RSDSFM_CONVERGENCE; summary->reason = RSDSFM_REASON_FUNCTION_TOL; return RSDSFM_OK; }
    const size_t nn = (size_t)n;
    RS_TRY(ensure(ctx, ctx->pix, sizeof(double2) * nn));
    RS_TRY(ensure(ctx, ctx->dA, sizeof(double) * nn));
    RS_TRY(ensure(ctx, ctx->dB, sizeof(double) * nn));
    RS_TRY(ensure(ctx, ctx->scale_e, sizeof(double) * nn));
    double2 *aa = (double2 *)ctx->pix.p;
    double *d0 = (double *)ctx->dA.p, *d1 = (double *)ctx->dB.p;
    const int grid = grid_for(ctx, n, 8);
    k_depth_gather<<<grid, kThreads, 0, ctx->stream>>>(alpha, alpha_k, n, aa, d0);
    ctx->launches++;
    bool finite = isfinite(k);
    for (int j = 0; j < 3; ++j) finite = finite && isfinite(v[j]) && isfinite(w[j]);
    double *dfin = nullptr;
    if (!finite) {
        summary->termination = RSDSFM_FAILURE; summary->reason = RSDSFM_REASON_NONFINITE_INPUT;
    } else {
        RefineData D{(const double2 *)coord, (const double2 *)flow, aa, (double *)ctx->scale_e.p, n};
        Motion mot;
        for (int j = 0; j < 3; ++j) { mot.v[j] = v[j]; mot.w[j] = w[j]; }
        mot.k = k;
        RS_TRY(lm_solve_device(ctx, D, d0, d1, 0, mot, o, summary, &dfin));
    }
    if (dfin == nullptr) {
        k_depth_gather<<<grid, kThreads, 0, ctx->stream>>>(alpha, alpha_k, n, aa, d0);
        ctx->launches++;
        dfin = d0;
    }
    k_copy<<<grid, kThreads, 0, ctx->stream>>>(dfin, n, inv_depth);
    ctx->launches++;
    return RSDSFM_OK;
}

}  // namespace rsdsfm
