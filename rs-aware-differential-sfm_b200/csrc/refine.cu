// refine.cu -- GPU Levenberg-Marquardt for the reference's two Ceres problems:
//   a8  nonlinear_refinement::estimateInverseDepths  (nonlinearRefinement.cc:109-180)  NF = 0
//   a9  nonlinear_refinement::nonLinearRefinement    (nonlinearRefinement.cc:183-252)  NF = 6 | 7
//
// ONE persistent cooperative kernel runs the whole solve: one CTA per SM stays resident and
// loops over LM phases; a phase is a coalesced pass over the residual blocks followed by a grid
// reduction (registers -> shared-memory transpose -> warp shuffles -> one row per CTA -> the last
// CTA to arrive sums the rows in fixed order) and the O(1) controller (lm_controller.h: Ceres 1.14
// trust-region semantics, 7x7 Cholesky) executed by that last CTA, which then releases the grid.
// No host round trip per iteration; the host launches once and reads one summary back.
//
//   pass A (evaluation at x): residual + analytic Jacobian, closed-form 1x1 Schur elimination of
//       the pixel's inverse depth, FP64 accumulation of the radius-independent factors G1, G2,
//       h1, h2 (two sparse rank-1 updates per pixel), cost, |x|^2, max depth gradient.
//   pass B (candidate): depth back-substitution, candidate point + cost, model cost change, |step|^2.
//   A rejected step re-solves from the stored factors: it costs a pass B only.
//
// Data layout in HBM (structure of arrays, one entry per residual block, 16-byte vector loads):
//   xy[m] double2 (x, y) | uu[m] double2 (ux, uy; Q1 pairing applied once by the gather kernel)
//   aa[m] double2 (alpha, alpha_k) | d[2][m] inverse depth ping-pong (x / candidate)
//   se[m] Jacobi scale of the depth column, fixed at iteration 0 (only used for the clamp test)
// Algorithmic bytes: pass A 24 B, pass B 32 B per residual block.
#include <cooperative_groups.h>

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

struct ExcEntry {            // a pixel whose LM diagonal is clamped (or whose e-column is degenerate)
    double ees, se2, er;     // s_e^2 e^Te, s_e^2, e^T r
    double fe[kMaxNF];       // F^T e
};

// Broadcast block: written by the controller CTA, read by every CTA at the start of a phase.
struct Bcast {
    int next, which_x, first, pad0;
    Motion mot, cand;
    double delta_f[kMaxNF];
    double radius;
};

// Device-resident control block of one solve.
struct LmShared {
    LmController ctl;
    Motion base;             // values of the motion parameters that are not free
    Bcast bc;
    // ---- per-phase device timing (globaltimer ns): [0] pass A total, [1] phases, [2] pass B total, [3] phases
    unsigned long long t_phase[4];
    // ---- grid synchronisation
    unsigned int arrive, generation;
    unsigned int n_exc, exc_overflow;
    int error;
    int nonfinite_input;     // LAST field: raised by the gather kernel, preserved by the control-block upload
};

// ------------------------------------------------------------------------------------------
// gather: API arrays -> SoA records.  Residual i pairs inlier i with flow(:, i) of the array the
// caller passed (reference behaviour, nonlinearRefinement.cc:209-212) or flow(:, flow_index[i]).
// ------------------------------------------------------------------------------------------
__global__ void k_refine_gather(const double *__restrict__ flow, const double *__restrict__ inliers3,
                                const double *__restrict__ alpha, const double *__restrict__ alpha_k,
                                const int32_t *__restrict__ flow_index, int m, double2 *xy, double2 *uu, double2 *aa,
                                double *d0, LmShared *sh)
{
    int bad = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const double x = inliers3[3 * (size_t)i], y = inliers3[3 * (size_t)i + 1], z = inliers3[3 * (size_t)i + 2];
        const int fi = flow_index ? flow_index[i] : i;
        const double2 u = reinterpret_cast<const double2 *>(flow)[fi];
        xy[i] = make_double2(x, y);
        uu[i] = u;
        aa[i] = make_double2(alpha[i], alpha_k[i]);
        const double d = 1.0 / z;                              // nonlinearRefinement.cc:213
        d0[i] = d;
        if (!isfinite(d)) bad = 1;
    }
    // solver.cc: non-finite initial parameter values => FAILURE before any evaluation
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(&sh->nonfinite_input, 1);
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

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp(double x)
{   // MUFU.RCP64H seed + two Newton steps: full double precision for normal, finite x
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double t = fma(-x, r, 1.0);
    r = fma(r, t, r);
    t = fma(-x, r, 1.0);
    r = fma(r, t, r);
    return r;
}

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct PhaseParams {           // shared-memory copy of the broadcast block (+ options)
    int next, which_x, first, pad0;
    Motion mot, cand;
    double delta_f[kMaxNF];
    double radius;
    double min_diag, max_diag;
    int error;
};
static_assert(offsetof(PhaseParams, radius) == offsetof(Bcast, radius), "PhaseParams must start with Bcast");

template <int NF> constexpr int red_rows()
{   // rows of the shared reduction scratch: pass A values, pass B values (5), exception sums (kTri + kMaxNF)
    int r = 2 + NF * (NF + 1) + 2 * NF + 2;
    if (r < 5) r = 5;
    if (NF > 0 && r < kTri + kMaxNF) r = kTri + kMaxNF;
    return r;
}
template <int NF> constexpr int kRedRows = red_rows<NF>();

struct Loaded {
    double2 p, u, a;
    double d, se;
};
__device__ __forceinline__ Loaded load_px(const RefineData &D, const double *__restrict__ d, int i, bool want_se)
{
    Loaded L;
    L.p = D.xy[i]; L.u = D.uu[i]; L.a = D.aa[i]; L.d = d[i];
    L.se = want_se ? D.se[i] : 1.0;
    return L;
}

// NV values per thread (first NS sums, then NM maxima) -> one row of NV doubles for this CTA.
// red: shared scratch of NV * kThreads doubles.  Fixed order => bit-reproducible.
template <int NS, int NM>
__device__ __forceinline__ void cta_reduce(const double (&v)[NS + NM], double *red, double *row)
{
    constexpr int NV = NS + NM;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int j = 0; j < NV; ++j) red[j * kThreads + tid] = v[j];
    __syncthreads();
    for (int j = warp; j < NV; j += kWarps) {
        const double *c = red + j * kThreads + lane;
        double s = c[0];
        if (j < NS) {
#pragma unroll
            for (int k = 1; k < kWarps; ++k) s += c[32 * k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        } else {
#pragma unroll
            for (int k = 1; k < kWarps; ++k) s = fmax(s, c[32 * k]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s = fmax(s, __shfl_xor_sync(0xffffffffu, s, o));
        }
        if (lane == 0) row[j] = s;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Pass A body: NS = 2 + 2*TRI + 2*NF sums (cost2, sum d^2, G1, G2, h1, h2), NM = 2 maxima
// ------------------------------------------------------------------------------------------
template <int NF>
struct PassA {
    static constexpr int TRI = NF * (NF + 1) / 2;
    static constexpr int oG1 = 2, oG2 = oG1 + TRI, oH1 = oG2 + TRI, oH2 = oH1 + NF;
    static constexpr int NS = oH2 + NF, NM = 2, NV = NS + NM;
    static constexpr int iGMAX = NS, iBAD = NS + 1;
};

// f = F^T (pi0, pi1)  with  F = -beta [d A | B | (dbeta/beta) p]  (see rs_math.cuh)
template <int NF>
__device__ __forceinline__ void ft_times(double beta, double dbeta, double d, double x, double y, double xy, double xx1,
                                         double yy1, double p0, double p1, double pi0, double pi1, double (&f)[NF > 0 ? NF : 1])
{
    const double P0 = beta * pi0, P1 = beta * pi1;
    f[0] = -d * P0;
    f[1] = -d * P1;
    f[2] = d * fma(x, P0, y * P1);
    f[3] = fma(xy, P0, yy1 * P1);
    f[4] = -fma(xx1, P0, xy * P1);
    f[5] = fma(y, P0, -x * P1);
    if (NF == 7) f[6] = -dbeta * fma(p0, pi0, p1 * pi1);
}

template <int NF>
__device__ __forceinline__ void pass_a_pixel(const Loaded &L, const PhaseParams &P, double c2, double (&acc)[PassA<NF>::NV],
                                             const RefineData &D, int i, LmShared *sh, ExcEntry *exc, unsigned int exc_cap)
{
    using A = PassA<NF>;
    const double x = L.p.x, y = L.p.y, d = L.d;
    const Motion &m = P.mot;
    const double ak = fma(m.k, L.a.y, L.a.x);
    const double beta = c2 * ak;
    const double a0 = fma(-x, m.v[2], m.v[0]), a1 = fma(-y, m.v[2], m.v[1]);
    const double xy = x * y, xx1 = fma(x, x, 1.0), yy1 = fma(y, y, 1.0);
    const double b0 = fma(-xy, m.w[0], fma(xx1, m.w[1], -y * m.w[2]));
    const double b1 = fma(-yy1, m.w[0], fma(xy, m.w[1], x * m.w[2]));
    const double p0 = fma(d, a0, b0), p1 = fma(d, a1, b1);
    const double r0 = fma(-beta, p0, L.u.x), r1 = fma(-beta, p1, L.u.y);
    const double e0 = -beta * a0, e1 = -beta * a1;
    const double ee = fma(e0, e0, e1 * e1);
    double se = L.se;
    if (P.first) { se = 1.0 / (1.0 + sqrt(ee)); D.se[i] = se; }      // jacobian_scaling_, fixed at iteration 0
    acc[0] = fma(r0, r0, fma(r1, r1, acc[0]));
    acc[1] = fma(d, d, acc[1]);
    const double re = fma(e0, r0, e1 * r1);                          // e^T r
    acc[A::iGMAX] = fmax(acc[A::iGMAX], fabs(re));
    double bad = bad_flag(r0 + r1 + ee);
    if (NF > 0) {
        const double dbeta = (NF == 7) ? c2 * fma(-ak, 0.5 * c2, L.a.y) : 0.0;
        const double ees = ee * se * se;
        if (ees >= P.min_diag && ees <= P.max_diag) {
            // unclamped: projector = (n n^T + e e^T/(radius+1)) / e^Te -- accumulate the two factors
            const double mu = fast_rcp(ee);
            const double rn = fma(-e1, r0, e0 * r1);                 // n^T r, n = (-e1, e0)
            double fn[NF > 0 ? NF : 1], fe[NF > 0 ? NF : 1];
            ft_times<NF>(beta, dbeta, d, x, y, xy, xx1, yy1, p0, p1, -e1, e0, fn);
            ft_times<NF>(beta, dbeta, d, x, y, xy, xx1, yy1, p0, p1, e0, e1, fe);
            int t = 0;
#pragma unroll
            for (int j = 0; j < NF; ++j) {
                const double gn = mu * fn[j], ge = mu * fe[j];
                acc[A::oH1 + j] = fma(gn, rn, acc[A::oH1 + j]);
                acc[A::oH2 + j] = fma(ge, re, acc[A::oH2 + j]);
#pragma unroll
                for (int c = j; c < NF; ++c, ++t) {
                    acc[A::oG1 + t] = fma(gn, fn[c], acc[A::oG1 + t]);
                    acc[A::oG2 + t] = fma(ge, fe[c], acc[A::oG2 + t]);
                }
            }
            bad += bad_flag(mu);
        } else {
            // clamped LM diagonal (|e| ~ 0) or non-finite: F^TF / F^Tr go to G1 / h1, the
            // radius-dependent term q (F^Te)(e^TF) is applied by the controller from the exception list
            double F0[NF > 0 ? NF : 1], F1[NF > 0 ? NF : 1], fe[NF > 0 ? NF : 1];
            ft_times<NF>(beta, dbeta, d, x, y, xy, xx1, yy1, p0, p1, 1.0, 0.0, F0);
            ft_times<NF>(beta, dbeta, d, x, y, xy, xx1, yy1, p0, p1, 0.0, 1.0, F1);
            int t = 0;
#pragma unroll
            for (int j = 0; j < NF; ++j) {
                fe[j] = fma(F0[j], e0, F1[j] * e1);
                bad += bad_flag(fe[j]);
                acc[A::oH1 + j] += fma(F0[j], r0, F1[j] * r1);
#pragma unroll
                for (int c = j; c < NF; ++c, ++t) acc[A::oG1 + t] += fma(F0[j], F0[c], F1[j] * F1[c]);
            }
            const unsigned int slot = atomicAdd(&sh->n_exc, 1u);
            if (slot < exc_cap) {
                ExcEntry E;
                E.ees = ees; E.se2 = se * se; E.er = re;
#pragma unroll
                for (int j = 0; j < kMaxNF; ++j) E.fe[j] = (j < NF) ? fe[j] : 0.0;
                exc[slot] = E;
            } else {
                sh->exc_overflow = 1u;
            }
        }
    }
    acc[A::iBAD] = fmax(acc[A::iBAD], bad);
}

// ------------------------------------------------------------------------------------------
// Pass B body: sums mcc, step^2, candidate cost2; maxima bad_step, bad_cand
// ------------------------------------------------------------------------------------------
template <int NF>
__device__ __forceinline__ void pass_b_pixel(const Loaded &L, const PhaseParams &P, double c2, double c2c, double rfac,
                                             double inv_radius, double (&acc)[5], double *__restrict__ d_cand, int i)
{
    const double x = L.p.x, y = L.p.y, d = L.d;
    const Motion &m = P.mot;
    const double ak = fma(m.k, L.a.y, L.a.x);
    const double beta = c2 * ak;
    const double a0 = fma(-x, m.v[2], m.v[0]), a1 = fma(-y, m.v[2], m.v[1]);
    const double xy = x * y, xx1 = fma(x, x, 1.0), yy1 = fma(y, y, 1.0);
    const double b0 = fma(-xy, m.w[0], fma(xx1, m.w[1], -y * m.w[2]));
    const double b1 = fma(-yy1, m.w[0], fma(xy, m.w[1], x * m.w[2]));
    const double p0 = fma(d, a0, b0), p1 = fma(d, a1, b1);
    const double r0 = fma(-beta, p0, L.u.x), r1 = fma(-beta, p1, L.u.y);
    const double e0 = -beta * a0, e1 = -beta * a1;
    const double ee = fma(e0, e0, e1 * e1);
    // q = s_e^2 / (s_e^2 e^Te + clamp(s_e^2 e^Te)/radius)
    const double se2 = L.se * L.se;
    const double ees = ee * se2;
    double q;
    if (ees >= P.min_diag && ees <= P.max_diag) q = fast_rcp(ee) * rfac;              // rfac = radius/(radius+1)
    else q = se2 / (ees + fmin(fmax(ees, P.min_diag), P.max_diag) * inv_radius);
    // F delta_f = -beta (d A dv + B dw) - dbeta p dk
    double m0 = 0.0, m1 = 0.0;
    if (NF >= 6) {
        const double *df = P.delta_f;
        const double da0 = fma(-x, df[2], df[0]), da1 = fma(-y, df[2], df[1]);
        const double db0 = fma(-xy, df[3], fma(xx1, df[4], -y * df[5]));
        const double db1 = fma(-yy1, df[3], fma(xy, df[4], x * df[5]));
        m0 = -beta * fma(d, da0, db0);
        m1 = -beta * fma(d, da1, db1);
        if (NF == 7) {
            const double dbk = c2 * fma(-ak, 0.5 * c2, L.a.y) * df[6];
            m0 = fma(-dbk, p0, m0);
            m1 = fma(-dbk, p1, m1);
        }
    }
    const double delta_e = -q * fma(e0, r0 + m0, e1 * (r1 + m1));
    m0 = fma(e0, delta_e, m0);
    m1 = fma(e1, delta_e, m1);                                                      // J delta
    acc[0] += fma(m0, fma(0.5, m0, r0), m1 * fma(0.5, m1, r1));
    const double dc = d + delta_e;
    const double dd = d - dc;
    acc[1] = fma(dd, dd, acc[1]);
    d_cand[i] = dc;
    // candidate residual
    const Motion &c = P.cand;
    const double akc = fma(c.k, L.a.y, L.a.x);
    const double betac = c2c * akc;
    const double ca0 = fma(-x, c.v[2], c.v[0]), ca1 = fma(-y, c.v[2], c.v[1]);
    const double cb0 = fma(-xy, c.w[0], fma(xx1, c.w[1], -y * c.w[2]));
    const double cb1 = fma(-yy1, c.w[0], fma(xy, c.w[1], x * c.w[2]));
    const double s0 = fma(-betac, fma(dc, ca0, cb0), L.u.x), s1 = fma(-betac, fma(dc, ca1, cb1), L.u.y);
    acc[2] = fma(s0, s0, fma(s1, s1, acc[2]));
    acc[3] = fmax(acc[3], bad_flag(delta_e));
    acc[4] = fmax(acc[4], bad_flag(s0 + s1));
}

// ------------------------------------------------------------------------------------------
// The persistent kernel
// ------------------------------------------------------------------------------------------
constexpr int kParts = 3;                         // row segments summed concurrently in the final reduce
constexpr unsigned long long kWatchdogNs = 4000000000ull;   // 4 s: a stuck grid barrier aborts the solve

template <int NF>
__global__ void __launch_bounds__(kThreads, 1)
k_lm_persistent(RefineData D, double *d0, double *d1, LmShared *sh, double *partials, ExcEntry *exc, unsigned int exc_cap,
                const double *z_in, int z_stride, double *out, int invert_out)
{
    using A = PassA<NF>;
    extern __shared__ double red[];                       // kRedRows<NF> * kThreads doubles
    __shared__ PhaseParams P;
    __shared__ LmController s_ctl;
    __shared__ double fin[kRedRows<NF>];
    __shared__ double part[kRedRows<NF> * kParts];
    __shared__ ExcSums s_exc;
    __shared__ int s_flag[4];                             // [0] is_last, [1] next, [2] n_exc

    const int tid = threadIdx.x;
    const int stride = gridDim.x * kThreads;
    const int start = blockIdx.x * kThreads + tid;
    unsigned int gen = 0;

    for (;;) {
        // ---- phase parameters
        if (tid < (int)(sizeof(Bcast) / sizeof(int)))
            reinterpret_cast<int *>(&P)[tid] = __ldcg(reinterpret_cast<const int *>(&sh->bc) + tid);
        if (tid == 64) {
            P.min_diag = __ldcg(&sh->ctl.opt.min_lm_diagonal); P.max_diag = __ldcg(&sh->ctl.opt.max_lm_diagonal);
            P.error = __ldcg(&sh->error);
        }
        __syncthreads();
        if (P.next == LM_DONE || P.error) break;
        const bool run_a = (P.next == LM_RUN_A);
        double *dx = P.which_x ? d1 : d0;
        double *dcand = P.which_x ? d0 : d1;
        double *row = partials + (size_t)blockIdx.x * A::NV;
        const unsigned long long t_begin = (blockIdx.x == 0 && tid == 0) ? globaltimer() : 0ull;

        if (run_a) {
            double acc[A::NV];
#pragma unroll
            for (int j = 0; j < A::NV; ++j) acc[j] = 0.0;
            const double c2 = 2.0 / (2.0 + P.mot.k);
            const bool want_se = !P.first;
            int i = start;
            Loaded cur;
            if (i < D.m) cur = load_px(D, dx, i, want_se);
            while (i < D.m) {
                const int ni = i + stride;
                Loaded nxt = cur;
                if (ni < D.m) nxt = load_px(D, dx, ni, want_se);         // software prefetch
                pass_a_pixel<NF>(cur, P, c2, acc, D, i, sh, exc, exc_cap);
                cur = nxt;
                i = ni;
            }
            cta_reduce<A::NS, A::NM>(acc, red, row);
        } else {
            double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
            const double c2 = 2.0 / (2.0 + P.mot.k), c2c = 2.0 / (2.0 + P.cand.k);
            const double rfac = P.radius / (P.radius + 1.0), inv_radius = 1.0 / P.radius;
            int i = start;
            Loaded cur;
            if (i < D.m) cur = load_px(D, dx, i, true);
            while (i < D.m) {
                const int ni = i + stride;
                Loaded nxt = cur;
                if (ni < D.m) nxt = load_px(D, dx, ni, true);
                pass_b_pixel<NF>(cur, P, c2, c2c, rfac, inv_radius, acc, dcand, i);
                cur = nxt;
                i = ni;
            }
            cta_reduce<3, 2>(acc, red, row);
        }

        // ---- grid barrier: the last CTA to arrive reduces the rows and runs the controller
        if (tid == 0) {
            __threadfence();
            const unsigned int ticket = atomicAdd(&sh->arrive, 1u);
            s_flag[0] = (ticket == (gen + 1u) * gridDim.x - 1u) ? 1 : 0;
        }
        __syncthreads();
        if (s_flag[0]) {
            __threadfence();
            const int nv = run_a ? A::NV : 5, ns = run_a ? A::NS : 3;
            const int rows = gridDim.x;
            for (int idx = tid; idx < nv * kParts; idx += kThreads) {
                const int j = idx / kParts, pt = idx - j * kParts;
                const int lo = (rows * pt) / kParts, hi = (rows * (pt + 1)) / kParts;
                const double *p = partials + j;
                double s = (j < ns) ? 0.0 : -INFINITY;
                if (j < ns) for (int b = lo; b < hi; ++b) s += __ldcg(p + (size_t)b * A::NV);
                else        for (int b = lo; b < hi; ++b) s = fmax(s, __ldcg(p + (size_t)b * A::NV));
                part[idx] = s;
            }
            // controller state: global -> shared
            for (int w = tid; w < (int)(sizeof(LmController) / sizeof(int)); w += kThreads)
                reinterpret_cast<int *>(&s_ctl)[w] = __ldcg(reinterpret_cast<const int *>(&sh->ctl) + w);
            __syncthreads();
            if (tid < nv) {
                double s = part[tid * kParts];
                if (tid < ns) for (int pt = 1; pt < kParts; ++pt) s += part[tid * kParts + pt];
                else          for (int pt = 1; pt < kParts; ++pt) s = fmax(s, part[tid * kParts + pt]);
                fin[tid] = s;
            }
            __syncthreads();
            if (tid == 0) {
                LmNext nx;
                if (run_a) {
                    EvalSums e;
                    e.cost = 0.5 * fin[0]; e.sumsq_d = fin[1]; e.gmax_e = fin[A::iGMAX]; e.bad = fin[A::iBAD];
                    for (int t = 0; t < kTri; ++t) { e.G1[t] = (t < A::TRI) ? fin[A::oG1 + t] : 0.0; e.G2[t] = (t < A::TRI) ? fin[A::oG2 + t] : 0.0; }
                    for (int j = 0; j < kMaxNF; ++j) { e.h1[j] = (j < NF) ? fin[A::oH1 + j] : 0.0; e.h2[j] = (j < NF) ? fin[A::oH2 + j] : 0.0; }
                    nx = s_ctl.on_eval(e);
                } else {
                    CandSums c;
                    c.mcc = fin[0]; c.step_sq = fin[1]; c.cand_cost = 0.5 * fin[2]; c.bad_step = fin[3]; c.bad_cand = fin[4];
                    nx = s_ctl.on_candidate(c);
                }
                s_flag[1] = (int)nx;
                unsigned int ne = __ldcg(&sh->n_exc);
                s_flag[2] = (int)(ne < exc_cap ? ne : exc_cap);
            }
            __syncthreads();
            // ---- (re)solve at the current radius; the clamped-pixel correction is summed by the whole CTA
            while (s_flag[1] == (int)LM_SOLVE) {
                const int ne = s_flag[2];
                if (NF > 0 && ne > 0) {
                    const double R = s_ctl.radius, lo = s_ctl.opt.min_lm_diagonal, hi = s_ctl.opt.max_lm_diagonal;
                    double a[kTri + kMaxNF];
#pragma unroll
                    for (int j = 0; j < kTri + kMaxNF; ++j) a[j] = 0.0;
                    for (int k = tid; k < ne; k += kThreads) {
                        ExcEntry E;
                        for (int w = 0; w < (int)(sizeof(ExcEntry) / sizeof(double)); ++w)
                            reinterpret_cast<double *>(&E)[w] = __ldcg(reinterpret_cast<const double *>(exc + k) + w);
                        const double q = E.se2 / (E.ees + fmin(fmax(E.ees, lo), hi) / R);
                        int t = 0;
#pragma unroll
                        for (int j = 0; j < NF; ++j) {
                            const double qf = q * E.fe[j];
                            a[kTri + j] = fma(qf, E.er, a[kTri + j]);
#pragma unroll
                            for (int c = j; c < NF; ++c, ++t) a[t] = fma(qf, E.fe[c], a[t]);
                        }
                    }
                    cta_reduce<kTri + kMaxNF, 0>(a, red, fin);
                    if (tid < kTri) s_exc.S[tid] = fin[tid];
                    if (tid < kMaxNF) s_exc.rhs[tid] = fin[kTri + tid];
                    __syncthreads();
                }
                if (tid == 0) s_flag[1] = (int)s_ctl.solve_step((NF > 0 && ne > 0) ? &s_exc : nullptr);
                __syncthreads();
            }
            // ---- publish the next phase
            if (tid == 0) {
                const LmNext nx = (LmNext)s_flag[1];
                if (s_ctl.accepted_last && nx == LM_RUN_A) sh->bc.which_x = P.which_x ^ 1;
                Motion mo = sh->base, ca = sh->base;
                if (NF >= 6) for (int j = 0; j < 3; ++j) {
                    mo.v[j] = s_ctl.f[j]; mo.w[j] = s_ctl.f[3 + j];
                    ca.v[j] = s_ctl.f[j] + s_ctl.delta_f[j]; ca.w[j] = s_ctl.f[3 + j] + s_ctl.delta_f[3 + j];
                }
                if (NF == 7) { mo.k = s_ctl.f[6]; ca.k = s_ctl.f[6] + s_ctl.delta_f[6]; }
                if (nx == LM_DONE && s_ctl.termination == RSDSFM_FAILURE) mo = sh->base;   // Ceres restores the start values
                sh->bc.mot = mo; sh->bc.cand = ca;
                for (int j = 0; j < kMaxNF; ++j) sh->bc.delta_f[j] = s_ctl.delta_f[j];
                sh->bc.radius = s_ctl.radius;
                sh->bc.first = 0;
                sh->bc.next = (int)nx;
                if (nx == LM_RUN_A) sh->n_exc = 0u;       // a new evaluation rebuilds the exception list
                if (t_begin) {
                    const unsigned long long dt = globaltimer() - t_begin;
                    sh->t_phase[run_a ? 0 : 2] += dt; sh->t_phase[run_a ? 1 : 3] += 1ull;
                }
            }
            for (int w = tid; w < (int)(sizeof(LmController) / sizeof(int)); w += kThreads)
                reinterpret_cast<int *>(&sh->ctl)[w] = reinterpret_cast<const int *>(&s_ctl)[w];
            __syncthreads();
            if (tid == 0) { __threadfence(); st_release(&sh->generation, gen + 1u); }
        }
        // ---- everybody waits for the controller's release
        if (tid == 0) {
            const unsigned long long t0 = globaltimer();
            while (ld_acquire(&sh->generation) <= gen) {
                __nanosleep(40);
                if (globaltimer() - t0 > kWatchdogNs) { sh->error = 1; break; }
            }
            if (t_begin && !s_flag[0]) {
                const unsigned long long dt = globaltimer() - t_begin;
                sh->t_phase[run_a ? 0 : 2] += dt; sh->t_phase[run_a ? 1 : 3] += 1ull;
            }
        }
        __syncthreads();
        gen++;
    }

    // ---- epilogue: write the result (z = 1/d for a9, d for a8).  On FAILURE Ceres restores the
    // start values (solver.cc Minimize): 1/z_in for a9 (double reciprocal, :213/:247), 1.0 for a8.
    const bool failed = (__ldcg(&sh->ctl.termination) == RSDSFM_FAILURE) || P.error;
    const double *dfin = P.which_x ? d1 : d0;
    for (int i = start; i < D.m; i += stride) {
        double dv;
        if (failed) dv = z_in ? 1.0 / z_in[(size_t)i * z_stride] : 1.0;
        else dv = dfin[i];
        out[i] = invert_out ? 1.0 / dv : dv;
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int NF>
static int launch_persistent(rsdsfm_ctx *ctx, RefineData D, double *d0, double *d1, LmShared *sh, double *partials,
                             ExcEntry *exc, unsigned int exc_cap, const double *z_in, int z_stride, double *out,
                             int invert_out, int grid)
{
    const size_t smem = sizeof(double) * (size_t)kRedRows<NF> * kThreads;
    RS_CUDA(ctx, cudaFuncSetAttribute(k_lm_persistent<NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void *args[] = {&D, &d0, &d1, &sh, &partials, &exc, &exc_cap, &z_in, &z_stride, &out, &invert_out};
    if (ctx->profile) cudaEventRecord(ctx->pe0, ctx->stream);
    RS_CUDA(ctx, cudaLaunchCooperativeKernel((void *)k_lm_persistent<NF>, dim3(grid), dim3(kThreads), args, smem, ctx->stream));
    if (ctx->profile) cudaEventRecord(ctx->pe1, ctx->stream);
    ctx->launches++;
    return RSDSFM_OK;
}

__global__ void k_apply_input_flag(LmShared *sh)
{   // "gather found a non-finite start depth" => the FAILURE Ceres reports before evaluating anything
    if (sh->nonfinite_input) {
        sh->bc.next = LM_DONE;
        sh->ctl.termination = RSDSFM_FAILURE;
        sh->ctl.reason = RSDSFM_REASON_NONFINITE_INPUT;
    }
}

// Queues one LM solve on the context's stream (no host synchronisation).  The control block
// (ctx->lm_shared) receives the result; lm_collect() reads it back.
static int lm_solve_async(rsdsfm_ctx *ctx, const RefineData &D, double *d0, double *d1, int nf, const Motion &mot0,
                          const rsdsfm_lm_options &opt, const double *z_in, int z_stride, double *out, int invert_out,
                          bool keep_input_flag)
{
    RS_TRY(ensure(ctx, ctx->lm_shared, sizeof(LmShared)));
    LmShared *sh = (LmShared *)ctx->lm_shared.p;
    const int grid = ctx->num_sms;                      // one persistent CTA per SM
    const int nv = (nf == 0) ? PassA<0>::NV : (nf == 6 ? PassA<6>::NV : PassA<7>::NV);
    RS_TRY(ensure(ctx, ctx->partials, sizeof(double) * (size_t)grid * nv));
    if (ctx->exc_cap < 4096) ctx->exc_cap = 4096;
    RS_TRY(ensure(ctx, ctx->exc, sizeof(ExcEntry) * (size_t)ctx->exc_cap));
    RS_TRY(ensure_pinned(ctx, sizeof(LmShared) * 2 + 1024));

    // initial control block, staged through pinned memory (second half; the first half receives results)
    LmShared *h = (LmShared *)((char *)ctx->pinned + sizeof(LmShared));
    memset(h, 0, sizeof(LmShared));
    double f0[kMaxNF] = {0, 0, 0, 0, 0, 0, 0};
    if (nf >= 6) { for (int j = 0; j < 3; ++j) { f0[j] = mot0.v[j]; f0[3 + j] = mot0.w[j]; } }
    if (nf == 7) f0[6] = mot0.k;
    h->ctl.init(opt, nf, f0);
    h->base = mot0; h->bc.mot = mot0; h->bc.cand = mot0;
    h->bc.next = LM_RUN_A; h->bc.which_x = 0; h->bc.first = 1;
    h->bc.radius = h->ctl.radius;
    bool finite = isfinite(mot0.k);
    for (int j = 0; j < 3; ++j) finite = finite && isfinite(mot0.v[j]) && isfinite(mot0.w[j]);
    if (!finite) {              // solver.cc: non-finite parameter blocks => FAILURE, nothing evaluated
        h->bc.next = LM_DONE; h->ctl.termination = RSDSFM_FAILURE; h->ctl.reason = RSDSFM_REASON_NONFINITE_INPUT;
    }
    if (keep_input_flag) {
        // nonfinite_input is the LAST field: upload everything before it, keep what the gather kernel raised
        RS_CUDA(ctx, cudaMemcpyAsync(sh, h, offsetof(LmShared, nonfinite_input), cudaMemcpyHostToDevice, ctx->stream));
        k_apply_input_flag<<<1, 1, 0, ctx->stream>>>(sh);
        ctx->launches++;
    } else {
        RS_CUDA(ctx, cudaMemcpyAsync(sh, h, sizeof(LmShared), cudaMemcpyHostToDevice, ctx->stream));
    }
    double *partials = (double *)ctx->partials.p;
    ExcEntry *exc = (ExcEntry *)ctx->exc.p;
    const unsigned int cap = (unsigned int)ctx->exc_cap;
    if (nf == 0) return launch_persistent<0>(ctx, D, d0, d1, sh, partials, exc, cap, z_in, z_stride, out, invert_out, grid);
    if (nf == 6) return launch_persistent<6>(ctx, D, d0, d1, sh, partials, exc, cap, z_in, z_stride, out, invert_out, grid);
    return launch_persistent<7>(ctx, D, d0, d1, sh, partials, exc, cap, z_in, z_stride, out, invert_out, grid);
}

// Reads the control block back (synchronises the stream).  RSDSFM_ERR_INTERNAL when the in-kernel
// watchdog tripped; *overflow: the exception list was too small (ctx->exc_cap was raised: retry).
int lm_collect(rsdsfm_ctx *ctx, int nf, int m, Motion *mot, rsdsfm_lm_summary *summary, bool *overflow)
{
    LmShared *sh = (LmShared *)ctx->lm_shared.p;
    LmShared *h = (LmShared *)ctx->pinned;
    RS_CUDA(ctx, cudaMemcpyAsync(h, sh, sizeof(LmShared), cudaMemcpyDeviceToHost, ctx->stream));
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h->error) return fail(ctx, RSDSFM_ERR_INTERNAL, "LM kernel: grid barrier watchdog tripped");
    *overflow = h->exc_overflow != 0;
    if (*overflow) { ctx->exc_cap = m + 1024; return RSDSFM_OK; }
    float kms = 0.f;
    if (ctx->profile) cudaEventElapsedTime(&kms, ctx->pe0, ctx->pe1);
    if (summary) {
        h->ctl.fill_summary(summary);
        summary->device_ms = (double)(h->t_phase[0] + h->t_phase[2]) * 1e-6;
    }
    if (ctx->profile) {
        ctx->prof[0] += (double)h->t_phase[0] * 1e-6; ctx->prof[1] += (double)h->t_phase[1];
        ctx->prof[2] += (double)h->t_phase[1] * m;
        ctx->prof[3] += (double)h->t_phase[2] * 1e-6; ctx->prof[4] += (double)h->t_phase[3];
        ctx->prof[5] += (double)h->t_phase[3] * m;
        ctx->prof[6] += kms; ctx->prof[7] += 1.0;
    }
    if (mot && h->ctl.termination != RSDSFM_FAILURE) {
        if (nf >= 6) for (int j = 0; j < 3; ++j) { mot->v[j] = h->ctl.f[j]; mot->w[j] = h->ctl.f[3 + j]; }
        if (nf == 7) mot->k = h->ctl.f[6];
    }
    return RSDSFM_OK;
}

// Device address of the refined motion (v[3], w[3], k) inside the control block, for the
// rectification stage that follows on the same stream.
const double *lm_motion_device(rsdsfm_ctx *ctx)
{
    return reinterpret_cast<const double *>((const char *)ctx->lm_shared.p + offsetof(LmShared, bc) + offsetof(Bcast, mot));
}

// a9 on device pointers: queues gather + solve on the stream, no synchronisation.
int refine_async(rsdsfm_ctx *ctx, const double *flow, const double *inliers3, const double *alpha,
                 const double *alpha_k, int m, const double *v, const double *w, double k, int const_acc,
                 const int32_t *flow_index, const rsdsfm_lm_options *opts, double *z_out)
{
    rsdsfm_lm_options o;
    if (opts) o = *opts; else rsdsfm_lm_default_options(&o);
    const size_t mm = (size_t)m;
    RS_TRY(ensure(ctx, ctx->pix, sizeof(double2) * 3 * mm));
    RS_TRY(ensure(ctx, ctx->dA, sizeof(double) * mm));
    RS_TRY(ensure(ctx, ctx->dB, sizeof(double) * mm));
    RS_TRY(ensure(ctx, ctx->scale_e, sizeof(double) * mm));
    RS_TRY(ensure(ctx, ctx->lm_shared, sizeof(LmShared)));
    double2 *xy = (double2 *)ctx->pix.p, *uu = xy + mm, *aa = uu + mm;
    double *d0 = (double *)ctx->dA.p, *d1 = (double *)ctx->dB.p;
    LmShared *sh = (LmShared *)ctx->lm_shared.p;
    RS_CUDA(ctx, cudaMemsetAsync(&sh->nonfinite_input, 0, sizeof(int), ctx->stream));
    k_refine_gather<<<grid_for(ctx, m, 8), kThreads, 0, ctx->stream>>>(flow, inliers3, alpha, alpha_k, flow_index, m, xy, uu,
                                                                      aa, d0, sh);
    ctx->launches++;
    RefineData D{xy, uu, aa, (double *)ctx->scale_e.p, m};
    Motion mot;
    for (int j = 0; j < 3; ++j) { mot.v[j] = v[j]; mot.w[j] = w[j]; }
    mot.k = k;
    return lm_solve_async(ctx, D, d0, d1, const_acc ? 7 : 6, mot, o, inliers3 + 2, 3, z_out, 1, true);
}

// a9, synchronous: returns the refined motion and the summary on the host.
int refine_device(rsdsfm_ctx *ctx, const double *flow, const double *inliers3, const double *alpha,
                  const double *alpha_k, int m, double *v, double *w, double *k, int const_acc,
                  const int32_t *flow_index, const rsdsfm_lm_options *opts, double *z_out, rsdsfm_lm_summary *summary)
{
    rsdsfm_lm_summary local;
    if (!summary) summary = &local;
    memset(summary, 0, sizeof *summary);
    if (m == 0) { summary->termination = RSDSFM_CONVERGENCE; summary->reason = RSDSFM_REASON_FUNCTION_TOL; return RSDSFM_OK; }
    for (int attempt = 0; attempt < 2; ++attempt) {
        RS_TRY(refine_async(ctx, flow, inliers3, alpha, alpha_k, m, v, w, *k, const_acc, flow_index, opts, z_out));
        Motion mot;
        for (int j = 0; j < 3; ++j) { mot.v[j] = v[j]; mot.w[j] = w[j]; }
        mot.k = *k;
        bool overflow = false;
        RS_TRY(lm_collect(ctx, const_acc ? 7 : 6, m, &mot, summary, &overflow));
        if (overflow) continue;                         // exception list enlarged: run again
        for (int j = 0; j < 3; ++j) { v[j] = mot.v[j]; w[j] = mot.w[j]; }
        *k = mot.k;
        return RSDSFM_OK;
    }
    return fail(ctx, RSDSFM_ERR_INTERNAL, "refine: exception list overflow persisted");
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
    k_depth_gather<<<grid_for(ctx, n, 8), kThreads, 0, ctx->stream>>>(alpha, alpha_k, n, aa, d0);
    ctx->launches++;
    RefineData D{(const double2 *)coord, (const double2 *)flow, aa, (double *)ctx->scale_e.p, n};
    Motion mot;
    for (int j = 0; j < 3; ++j) { mot.v[j] = v[j]; mot.w[j] = w[j]; }
    mot.k = k;
    RS_TRY(lm_solve_async(ctx, D, d0, d1, 0, mot, o, nullptr, 1, inv_depth, 0, false));
    bool overflow = false;
    return lm_collect(ctx, 0, n, nullptr, summary, &overflow);
}

}  // namespace rsdsfm
