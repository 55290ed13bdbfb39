// refine.cu -- GPU Levenberg-Marquardt for the reference's two Ceres problems:
//   a8  nonlinear_refinement::estimateInverseDepths  (nonlinearRefinement.cc:109-180)  NF = 0
//   a9  nonlinear_refinement::nonLinearRefinement    (nonlinearRefinement.cc:183-252)  NF = 6 | 7
//
// ONE persistent cooperative kernel runs the whole solve: one CTA per SM stays resident and
// loops over LM phases.  A phase is a sweep over the residual blocks, streamed through a ring of
// shared-memory stages by TMA bulk copies (cp.async.bulk + mbarrier complete_tx; one elected
// thread issues, kStages tiles of 512 blocks in flight per SM), followed by a grid reduction
// (shuffles -> one row per CTA -> the last CTA to arrive sums the rows in a fixed order) and the
// O(1) controller (lm_controller.h: Ceres 1.14 trust-region semantics; warp-parallel bookkeeping
// and register/shuffle Cholesky in ctl_on_eval / ctl_solve) run by that last CTA, which then
// releases the grid.  No host round trip per iteration; the host launches once and the kernel's
// last act is a summary the host reads back.
//
//   INIT phase (once): residual + analytic Jacobian at the start point, closed-form 1x1 Schur
//       elimination of the pixel's inverse depth, FP64 accumulation of the radius-independent
//       factors G1, G2, h1, h2 (two sparse rank-1 updates per pixel), cost, |x|^2, max gradient.
//   FUSED phase (one per LM iteration): depth back-substitution of the candidate step at x
//       (step_pair: candidate depth, model cost change, |step|^2) and, in the same sweep, the
//       evaluation of the next iteration's sums AT THE CANDIDATE (eval_pair), speculatively.
//   A rejected step re-solves from the stored factors at the smaller radius: no sweep at all.
//
// Data layout in HBM: tile-blocked structure of arrays, tile = 512 residual blocks:
//   blk[tile] = { xy[512] (x, y) | uu[512] (ux, uy; Q1 pairing applied once by the gather kernel)
//                 | aa[512] (alpha, alpha_k) } as double2, 24 KB contiguous = ONE bulk copy,
//   d[2][tiles*512] inverse depth ping-pong (x / candidate), 4 KB per tile.
// The Jacobi scale of a depth column (fixed at iteration 0 by Ceres) only matters for the
// min/max_lm_diagonal clamp; it is bounded from below by a global quantity, so the sweeps test
// e^Te against that bound and the handful of pixels below it (focus of expansion) go to an
// exception list that the controller CTA handles exactly (sorted by block index: reproducible).
// Traffic per residual block: INIT reads 56 B, FUSED reads 56 B and writes 8 B (algorithmic
// minimum, SURVEY.md 8d: 24 B / 56 B -- x, y, alpha, alpha_k are inputs of the C ABI).
#include "common.cuh"
#include "lm_controller.h"
#include "lm_layout.h"
#include "rs_math.cuh"

namespace rsdsfm {

// Tile-blocked SoA: tile t holds xy[kTile], uu[kTile], aa[kTile] (double2 each) contiguously, so a
// whole tile arrives with ONE TMA bulk copy (the per-copy issue cost, not bandwidth, limits small
// copies: measured 6 TB/s with 14 KB copies vs 16 TB/s from L2 / 7 TB/s from HBM with 32 KB copies).
struct RefineData {
    const double2 *blk;      // [num_tiles][3][kTile]
    int m;
};

struct ExcEntry {            // a pixel whose LM diagonal is (possibly) clamped, or whose e-column is degenerate:
    double ees, se2;         // s_e^2 e^Te, s_e^2     (the whole pixel is handled by the controller CTA)
    double r0, r1, e0, e1;   // residual and depth column
    double F0[kMaxNF], F1[kMaxNF];   // the two Jacobian rows of the free motion parameters
    double key;              // residual-block index: the controller sums the list in ascending key order, so
                             // the result does not depend on the order the atomics handed out the slots
};

// Broadcast block: written by the controller CTA, read by every CTA at the start of a phase.
struct Bcast {
    int next, which_x, first, cur_list;   // cur_list: which exception list belongs to the current point
    Motion mot, cand;
    double delta_f[kMaxNF];
    double radius;
    double ee_fast_min;
};

// Device-resident control block of one solve.
struct LmShared {
    LmController ctl;
    Motion base;             // start values of the motion (non-free parameters keep them)
    Bcast bc;
    // device timing (globaltimer ns): [0] pass A total, [1] phases, [2] pass B total, [3] phases,
    // [4..6] pass A pixel loop / CTA reduce / controller, [7..9] same for pass B
    unsigned long long t_phase[12];
    // ---- grid synchronisation
    unsigned int arrive, generation;
    unsigned int n_exc[4], exc_overflow, pad1;   // exception lists (k_lm_persistent: 2, k_lm_solve: 3 rotating: current / speculative / being cleared)
    int error;
    int nonfinite_input;     // LAST field: raised by the gather kernel, preserved by the control-block upload
};

// ------------------------------------------------------------------------------------------
// gather: API arrays -> SoA records.  Residual i pairs inlier i with flow(:, i) of the array the
// caller passed (reference behaviour, nonlinearRefinement.cc:209-212) or flow(:, flow_index[i]).
// ------------------------------------------------------------------------------------------
__global__ void k_refine_gather(const double *__restrict__ flow, const double *__restrict__ inliers3,
                                const double *__restrict__ alpha, const double *__restrict__ alpha_k,
                                const int32_t *__restrict__ flow_index, int m, double2 *blk, double *d0, LmShared *sh)
{
    int bad = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const double x = inliers3[3 * (size_t)i], y = inliers3[3 * (size_t)i + 1], z = inliers3[3 * (size_t)i + 2];
        const int fi = flow_index ? flow_index[i] : i;
        const double2 u = reinterpret_cast<const double2 *>(flow)[fi];
        blk[blk_index(i, 0)] = make_double2(x, y);
        blk[blk_index(i, 1)] = u;
        blk[blk_index(i, 2)] = make_double2(alpha[i], alpha_k[i]);
        const double d = 1.0 / z;                              // nonlinearRefinement.cc:213
        d0[i] = d;
        if (!isfinite(d)) bad = 1;
    }
    // solver.cc: non-finite initial parameter values => FAILURE before any evaluation
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(&sh->nonfinite_input, 1);
}

// a8 variant: depth starts at 1.0 (:140); coord / flow are copied so the solver's 16-byte
// alignment requirement never leaks into the C ABI.
__global__ void k_depth_gather(const double *__restrict__ coord, const double *__restrict__ flow,
                               const double *__restrict__ alpha, const double *__restrict__ alpha_k, int n, double2 *blk,
                               double *d0)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        blk[blk_index(i, 0)] = make_double2(coord[2 * (size_t)i], coord[2 * (size_t)i + 1]);
        blk[blk_index(i, 1)] = make_double2(flow[2 * (size_t)i], flow[2 * (size_t)i + 1]);
        blk[blk_index(i, 2)] = make_double2(alpha[i], alpha_k[i]);
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
__device__ __forceinline__ double fast_rsqrt(double x)
{   // MUFU.RSQ64H seed + two Newton steps (x > 0, normal)
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double hx = 0.5 * x;
    r = r * fma(-hx * r, r, 1.5);
    r = r * fma(-hx * r, r, 1.5);
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
// ---- TMA bulk copy + mbarrier (sm_90+ PTX; SASS: UBLKCP / SYNCS)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    const uint32_t a = smem_u32(bar);
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

struct PhaseParams {           // shared-memory copy of the broadcast block (+ options, start point)
    int next, which_x, first, cur_list;
    Motion mot, cand;
    double delta_f[kMaxNF];
    double radius;
    double ee_fast_min;
    double min_diag, max_diag;
    Motion base;
    int error;
};
static_assert(offsetof(PhaseParams, ee_fast_min) == offsetof(Bcast, ee_fast_min), "PhaseParams must start with Bcast");

constexpr int kStages = 7;                      // tiles in flight per CTA (7 x 28 KB of the 227 KB shared memory)
struct Stage {
    double2 xy[kTile], uu[kTile], aa[kTile];
    double d[kTile];
};
static_assert(sizeof(Stage) == 28672 && kTile == 2 * kThreads, "stage layout");

struct Loaded {
    double2 p, u, a;
    double d;
};

// ------------------------------------------------------------------------------------------
// Per-thread accumulator layout.  Sums: [0] sum r^2 and [1] sum d^2 at the evaluation point,
// G1, G2 (packed upper triangles), h1, h2, then the candidate-step sums mcc and |step|^2.
// Maxima (all >= 0): max|e^T r|, bad evaluation, max e^Te, bad step, bad residual.
// ------------------------------------------------------------------------------------------
template <int NF>
struct Acc {
    static constexpr int TRI = NF * (NF + 1) / 2;
    static constexpr int oG1 = 2, oG2 = oG1 + TRI, oH1 = oG2 + TRI, oH2 = oH1 + NF, oMCC = oH2 + NF, oSTEP = oMCC + 1;
    static constexpr int NS = oSTEP + 1, NM = 5, NV = NS + NM;
    static constexpr int iGMAX = NS, iBAD = NS + 1, iEEMAX = NS + 2, iBADSTEP = NS + 3, iBADRES = NS + 4;
};
constexpr int kExcVals = kTri + kMaxNF;          // exception sums: S triangle + rhs
template <int NF> constexpr int kRowVals = (Acc<NF>::NV > kExcVals) ? Acc<NF>::NV : kExcVals;
static_assert(kRowVals<7> <= 96, "the final reduce covers three 32-lane column chunks");

// Per-THREAD accumulators.  The two Schur factors are split across lane pairs: even lanes keep
// K = G1 / H = h1 (the n-direction), odd lanes keep K = G2 / H = h2 (the e-direction); partners
// swap the half they do not keep with one shuffle per value.  This halves the accumulator
// registers (35 instead of 70 doubles for NF = 7), which is what lets two pixels per thread and the
// fused candidate+evaluation pass run without spilling.
template <int NF>
struct TAcc {
    static constexpr int TRI = NF * (NF + 1) / 2;
    static constexpr int oK = 2, oH = oK + TRI, oMCC = oH + NF, oSTEP = oMCC + 1;
    static constexpr int NS = oSTEP + 1, NM = 5, NV = NS + NM;
    static constexpr int iGMAX = NS, iBAD = NS + 1, iEEMAX = NS + 2, iBADSTEP = NS + 3, iBADRES = NS + 4;
};

// Per-thread accumulators -> one CTA row in the Acc<NF> layout (G1, G2, h1, h2 separated again).
// Butterfly shuffles inside each warp (K/H entries only over lanes of the same role), then the 8
// warp results are combined in warp order.  Fixed order => bit-reproducible.
template <int NF, int LD>
__device__ __forceinline__ void cta_reduce_roles(const double (&v)[TAcc<NF>::NV], double (*wpart)[LD], double *row)
{
    using T = TAcc<NF>;
    using A = Acc<NF>;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int j = 0; j < T::NV; ++j) {
        double x = v[j];
        const bool role_split = (j >= T::oK && j < T::oMCC);
        if (j < T::NS) {
#pragma unroll
            for (int o = 16; o > 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (!role_split) x += __shfl_xor_sync(0xffffffffu, x, 1);
        } else {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
        }
        if (role_split) {
            // lane 0: n-direction (G1 / h1), lane 1: e-direction (G2 / h2)
            const int jj = j - T::oK;
            const int dst = (jj < T::TRI) ? ((lane == 0 ? A::oG1 : A::oG2) + jj) : ((lane == 0 ? A::oH1 : A::oH2) + (jj - T::TRI));
            if (lane < 2) wpart[warp][dst] = x;
        } else if (lane == 0) {
            const int dst = (j < T::oK) ? j : (j < T::NS ? (A::oMCC + (j - T::oMCC)) : (A::NS + (j - T::NS)));
            wpart[warp][dst] = x;
        }
    }
    __syncthreads();
    if (tid < A::NV) {
        double x = wpart[0][tid];
        if (tid < A::NS) { for (int w = 1; w < kWarps; ++w) x += wpart[w][tid]; }
        else             { for (int w = 1; w < kWarps; ++w) x = fmax(x, wpart[w][tid]); }
        row[tid] = x;
    }
    __syncthreads();
}

// plain variant (all values reduced over all lanes): used for the exception sums
template <int NS, int LD>
__device__ __forceinline__ void cta_reduce_sums(const double (&v)[NS], double (*wpart)[LD], double *row)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        double x = v[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) wpart[warp][j] = x;
    }
    __syncthreads();
    if (tid < NS) {
        double x = wpart[0][tid];
        for (int w = 1; w < kWarps; ++w) x += wpart[w][tid];
        row[tid] = x;
    }
    __syncthreads();
}

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

// Jacobi scale of this pixel's depth column: 1/(1+|e(x0)|), e(x0) evaluated at the start motion.
__device__ __forceinline__ double depth_scale_at_start(const Loaded &L, const Motion &b)
{
    const double beta0 = (2.0 / (2.0 + b.k)) * fma(b.k, L.a.y, L.a.x);
    const double s0 = beta0 * fma(-L.p.x, b.v[2], b.v[0]), s1 = beta0 * fma(-L.p.y, b.v[2], b.v[1]);
    return 1.0 / (1.0 + sqrt(fma(s0, s0, s1 * s1)));
}

// Common per-pixel quantities at the point (m, d).
struct PxEval {
    double x, y, xy, xx1, yy1, ak, beta, p0, p1, r0, r1, e0, e1, ee;
};
__device__ __forceinline__ void px_eval(const Loaded &L, const Motion &m, double c2, double d, PxEval &E)
{
    E.x = L.p.x; E.y = L.p.y;
    E.ak = fma(m.k, L.a.y, L.a.x);
    E.beta = c2 * E.ak;
    const double a0 = fma(-E.x, m.v[2], m.v[0]), a1 = fma(-E.y, m.v[2], m.v[1]);
    E.xy = E.x * E.y; E.xx1 = fma(E.x, E.x, 1.0); E.yy1 = fma(E.y, E.y, 1.0);
    const double b0 = fma(-E.xy, m.w[0], fma(E.xx1, m.w[1], -E.y * m.w[2]));
    const double b1 = fma(-E.yy1, m.w[0], fma(E.xy, m.w[1], E.x * m.w[2]));
    E.p0 = fma(d, a0, b0); E.p1 = fma(d, a1, b1);
    E.r0 = fma(-E.beta, E.p0, L.u.x); E.r1 = fma(-E.beta, E.p1, L.u.y);
    E.e0 = -E.beta * a0; E.e1 = -E.beta * a1;
    E.ee = fma(E.e0, E.e0, E.e1 * E.e1);
}

// Rare path of the evaluation: a pixel whose LM diagonal may be clamped (|e| ~ 0, focus of
// expansion) or whose values are not finite.  Nothing of its Jacobian is accumulated by the
// thread: the pixel is listed and the controller CTA adds F^TF, F^Tr (radius independent) and
// subtracts q (F^Te)(e^TF), q (F^Te)(e^Tr) (radius dependent) itself.
template <int NF>
__device__ __noinline__ void eval_slow(const Loaded &L, double d, int index, const Motion &mot, double c2, bool first,
                                       const Motion &base, unsigned int *n_exc, unsigned int *overflow, ExcEntry *exc,
                                       unsigned int exc_cap)
{
    PxEval E;
    px_eval(L, mot, c2, d, E);
    const double dbeta = (NF == 7) ? c2 * fma(-E.ak, 0.5 * c2, L.a.y) : 0.0;
    const double se = first ? 1.0 / (1.0 + sqrt(E.ee)) : depth_scale_at_start(L, base);
    double F0[NF > 0 ? NF : 1], F1[NF > 0 ? NF : 1];
    ft_times<NF>(E.beta, dbeta, d, E.x, E.y, E.xy, E.xx1, E.yy1, E.p0, E.p1, 1.0, 0.0, F0);
    ft_times<NF>(E.beta, dbeta, d, E.x, E.y, E.xy, E.xx1, E.yy1, E.p0, E.p1, 0.0, 1.0, F1);
    const unsigned int slot = atomicAdd(n_exc, 1u);
    if (slot < exc_cap) {
        ExcEntry X;
        X.ees = E.ee * se * se; X.se2 = se * se;
        X.r0 = E.r0; X.r1 = E.r1; X.e0 = E.e0; X.e1 = E.e1; X.key = (double)index;
#pragma unroll
        for (int j = 0; j < kMaxNF; ++j) { X.F0[j] = (j < NF) ? F0[j] : 0.0; X.F1[j] = (j < NF) ? F1[j] : 0.0; }
        exc[slot] = X;
    } else {
        *overflow = 1u;
    }
}

// Evaluation (residual, Jacobian, Schur factors) of TWO residual blocks per thread at the point
// (mot, d[p]).  Branch-free on the common path; the lane pair (l, l^1) shares the rank-1 updates:
// the even lane applies the n-direction update of both lanes' pixels, the odd lane the e-direction.
template <int NF>
__device__ __forceinline__ void eval_pair(const Loaded (&L)[2], const double (&dv)[2], const bool (&valid)[2],
                                          int base_i, const Motion &mot,
                                          double c2, bool first, const PhaseParams &P, double (&acc)[TAcc<NF>::NV],
                                          unsigned int *n_exc, unsigned int *overflow, ExcEntry *exc, unsigned int exc_cap)
{
    using T = TAcc<NF>;
    const bool e_role = (threadIdx.x & 1) != 0;
    PxEval E[2];
    double d[2], mu[2];
    bool slow[2] = {false, false};
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        d[p] = dv[p];
        px_eval(L[p], mot, c2, d[p], E[p]);
        if (!valid[p]) { E[p].r0 = 0.0; E[p].r1 = 0.0; E[p].e0 = 0.0; E[p].e1 = 0.0; E[p].ee = 0.0; d[p] = 0.0; }
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        acc[0] = fma(E[p].r0, E[p].r0, fma(E[p].r1, E[p].r1, acc[0]));
        acc[1] = fma(d[p], d[p], acc[1]);
        const double re = fma(E[p].e0, E[p].r0, E[p].e1 * E[p].r1);     // e^T r
        acc[T::iGMAX] = fmax(acc[T::iGMAX], fabs(re));
        acc[T::iEEMAX] = fmax(acc[T::iEEMAX], E[p].ee);
        const double br = bad_flag(E[p].r0 + E[p].r1);
        acc[T::iBADRES] = fmax(acc[T::iBADRES], br);
        acc[T::iBAD] = fmax(acc[T::iBAD], br + bad_flag(E[p].ee));
        // is the LM diagonal of this depth certainly not clamped?  (first evaluation: the Jacobi
        // scale is 1/(1+|e|) of this very point; later: global lower bound of the scales)
        bool fast;
        if (first) {
            const double se = 1.0 / (1.0 + sqrt(E[p].ee));
            const double ees = E[p].ee * se * se;
            fast = (ees >= P.min_diag && ees <= P.max_diag);
        } else {
            fast = (E[p].ee >= P.ee_fast_min && E[p].ee <= P.max_diag);
        }
        slow[p] = (NF > 0) && valid[p] && !fast;
        const double r = fast_rsqrt(E[p].ee);                           // sqrt(mu) = 1/|e|
        mu[p] = (valid[p] && fast) ? r : 0.0;
    }
    if (NF > 0) {
        // projector = (n n^T + e e^T/(radius+1)) / e^Te: two radius-independent rank-1 factors.  Each lane
        // projects its pixel onto ITS direction (kept) and onto the partner's direction (handed over), both
        // pre-scaled by 1/|e| so that the accumulation is a plain symmetric rank-1 update.
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const double dbeta = (NF == 7) ? c2 * fma(-E[p].ak, 0.5 * c2, L[p].a.y) : 0.0;
            const double n0 = -E[p].e1, n1 = E[p].e0;                    // n = (-e1, e0)
            const double my0 = mu[p] * (e_role ? E[p].e0 : n0), my1 = mu[p] * (e_role ? E[p].e1 : n1);
            const double ot0 = mu[p] * (e_role ? n0 : E[p].e0), ot1 = mu[p] * (e_role ? n1 : E[p].e1);
            double kv[NF > 0 ? NF : 1], sv[NF > 0 ? NF : 1], pv[NF > 0 ? NF : 1];
            ft_times<NF>(E[p].beta, dbeta, d[p], E[p].x, E[p].y, E[p].xy, E[p].xx1, E[p].yy1, E[p].p0, E[p].p1, my0, my1, kv);
            ft_times<NF>(E[p].beta, dbeta, d[p], E[p].x, E[p].y, E[p].xy, E[p].xx1, E[p].yy1, E[p].p0, E[p].p1, ot0, ot1, sv);
            const double ks = fma(my0, E[p].r0, my1 * E[p].r1);
            const double ps = __shfl_xor_sync(0xffffffffu, fma(ot0, E[p].r0, ot1 * E[p].r1), 1);
#pragma unroll
            for (int j = 0; j < NF; ++j) pv[j] = __shfl_xor_sync(0xffffffffu, sv[j], 1);
            int t = 0;
#pragma unroll
            for (int j = 0; j < NF; ++j) {
                acc[T::oH + j] = fma(kv[j], ks, fma(pv[j], ps, acc[T::oH + j]));
#pragma unroll
                for (int c = j; c < NF; ++c, ++t) acc[T::oK + t] = fma(kv[j], kv[c], fma(pv[j], pv[c], acc[T::oK + t]));
            }
        }
        if (slow[0] || slow[1]) {
#pragma unroll
            for (int p = 0; p < 2; ++p)
                if (slow[p]) {
                    // the out-of-line call takes the pixel by address: hand it a copy made HERE, so that the
                    // common path keeps L[] in registers instead of spilling it to the stack every iteration
                    const Loaded Lc = L[p];
                    eval_slow<NF>(Lc, dv[p], base_i + (int)threadIdx.x + p * kThreads, mot, c2, first, P.base, n_exc, overflow, exc, exc_cap);
                }
        }
    }
}

// Candidate step of two residual blocks at the current point: depth back-substitution
// delta_d = -q e^T (r + F delta_f), model cost change, |step|^2; returns the candidate depths.
template <int NF>
__device__ __forceinline__ void step_pair(const Loaded (&L)[2], const bool (&valid)[2], const int (&idx)[2], const PhaseParams &P,
                                          double c2, double rfac, double inv_radius, double (&acc)[TAcc<NF>::NV],
                                          double *__restrict__ d_cand, double (&dc)[2])
{
    using A = TAcc<NF>;
    PxEval E[2];
    double q[2], m0[2], m1[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) px_eval(L[p], P.mot, c2, L[p].d, E[p]);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        // q = s_e^2 / (s_e^2 e^Te + clamp(s_e^2 e^Te)/radius)  ( = radius/((radius+1) e^Te) when not clamped )
        q[p] = fast_rcp(E[p].ee) * rfac;
        if (!(E[p].ee >= P.ee_fast_min && E[p].ee <= P.max_diag)) {
            const double se = depth_scale_at_start(L[p], P.base);
            const double se2 = se * se, ees = E[p].ee * se2;
            q[p] = se2 / (ees + fmin(fmax(ees, P.min_diag), P.max_diag) * inv_radius);
        }
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        // F delta_f = -beta (d A dv + B dw) - dbeta p dk
        m0[p] = 0.0; m1[p] = 0.0;
        if (NF >= 6) {
            const double *df = P.delta_f;
            const double da0 = fma(-E[p].x, df[2], df[0]), da1 = fma(-E[p].y, df[2], df[1]);
            const double db0 = fma(-E[p].xy, df[3], fma(E[p].xx1, df[4], -E[p].y * df[5]));
            const double db1 = fma(-E[p].yy1, df[3], fma(E[p].xy, df[4], E[p].x * df[5]));
            m0[p] = -E[p].beta * fma(L[p].d, da0, db0);
            m1[p] = -E[p].beta * fma(L[p].d, da1, db1);
            if (NF == 7) {
                const double dbk = c2 * fma(-E[p].ak, 0.5 * c2, L[p].a.y) * df[6];
                m0[p] = fma(-dbk, E[p].p0, m0[p]);
                m1[p] = fma(-dbk, E[p].p1, m1[p]);
            }
        }
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const double delta_e = -q[p] * fma(E[p].e0, E[p].r0 + m0[p], E[p].e1 * (E[p].r1 + m1[p]));
        const double j0 = fma(E[p].e0, delta_e, m0[p]), j1 = fma(E[p].e1, delta_e, m1[p]);      // J delta
        dc[p] = L[p].d + delta_e;
        const double dd = L[p].d - dc[p];
        if (valid[p]) {
            d_cand[idx[p]] = dc[p];
            acc[A::oMCC] += fma(j0, fma(0.5, j0, E[p].r0), j1 * fma(0.5, j1, E[p].r1));
            acc[A::oSTEP] = fma(dd, dd, acc[A::oSTEP]);
            acc[A::iBADSTEP] = fmax(acc[A::iBADSTEP], bad_flag(delta_e));
        } else {
            dc[p] = 1.0;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Warp-parallel controller steps.  The controller runs once per phase on the last CTA while the
// whole grid waits, and its code is cold in the instruction cache every time (the pixel loops
// evict it), so it is written as SMALL rolled loops executed by one warp on shared-memory data
// (lanes work on matrix entries side by side) instead of a long unrolled scalar sequence.
// Same arithmetic as LmController::on_eval_stored / solve_step (which the host-stepped RANSAC
// solver keeps using for its f-block-free problems).
// ------------------------------------------------------------------------------------------
__constant__ unsigned char kLowR[28] = {0, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6, 6};
__constant__ unsigned char kLowC[28] = {0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 4, 0, 1, 2, 3, 4, 5, 0, 1, 2, 3, 4, 5, 6};

template <int NF>
__device__ __noinline__ int ctl_on_eval(LmController &c)
{
    const int lane = threadIdx.x & 31;
    const bool bad = c.ev.bad > 0.0;
    int done = -1;
    if (c.phase == 0) {
        // IterationZero: the Jacobi scaling is fixed here
        if (bad) { if (lane == 0) { c.initial_cost = 0.0; c.finish(RSDSFM_FAILURE, RSDSFM_REASON_EVAL_FAILED); } done = LM_DONE; }
        else {
            if (lane < NF) {
                const int t = lane * NF - (lane * (lane - 1)) / 2;
                c.scale_f[lane] = 1.0 / (1.0 + sqrt(c.ev.G1[t] + c.ev.G2[t]));
            }
            if (lane == 0) {
                c.x_cost = c.ev.cost; c.initial_cost = c.ev.cost;
                const double t = 1.0 + sqrt(c.ev.ee_max);
                c.ee_fast_min = c.opt.min_lm_diagonal * t * t;
                c.step_is_successful = 1;            // IterationZero counts as a successful step (lm_controller.h)
            }
        }
    } else {
        // HandleSuccessfulStep: the evaluation at the new x
        if (bad) { if (lane == 0) c.finish(RSDSFM_FAILURE, RSDSFM_REASON_EVAL_FAILED); done = LM_DONE; }
        else if (lane == 0) { c.x_cost = c.ev.cost; c.step_is_successful = 1; }
    }
    __syncwarp();
    if (done >= 0) return done;
    // gradient max norm |x - Plus(x, -g)|_inf and |x|
    double g = 0.0, xs = 0.0;
    if (lane < NF) {
        const double f = c.f[lane];
        const double proj = f + (-(c.ev.h1[lane] + c.ev.h2[lane]));
        g = fabs(f - proj);
        xs = f * f;
    }
    for (int o = 16; o > 0; o >>= 1) { g = fmax(g, __shfl_xor_sync(0xffffffffu, g, o)); xs += __shfl_xor_sync(0xffffffffu, xs, o); }
    int nx = 0;
    if (lane == 0) {
        c.gmax = fmax(c.ev.gmax_e, g);
        c.x_norm = sqrt(c.ev.sumsq_d + xs);
        nx = (int)c.begin_iteration();
    }
    return __shfl_sync(0xffffffffu, nx, 0);
}

// LevenbergMarquardtStrategy::ComputeStep on the Schur-reduced system at the current radius.
// One warp: lane i keeps row i of the lower triangle in registers; pivots / multipliers travel by
// shuffle; the factor's columns are fetched once through the shared scratch Lm for the backward
// substitution.  Eigen::LLT semantics: the solve fails on a non-positive or NaN pivot.
template <int NF>
__device__ __noinline__ int ctl_solve(LmController &c, const ExcSums *exc, double (*Lm)[8], double *y_unused)
{
    (void)y_unused;
    const int lane = threadIdx.x & 31;
    int nx = 0;
    if (NF == 0) {
        if (lane == 0) { c.reuse_diagonal = 1; nx = (int)LM_RUN_B; }
        return __shfl_sync(0xffffffffu, nx, 0);
    }
    constexpr int N = NF > 0 ? NF : 1;
    const int i = lane < N ? lane : N - 1;                       // lanes >= N shadow the last row (results unused)
    const double radius = c.radius;
    const double sci = c.scale_f[i];
    if (!c.reuse_diagonal && lane < N) {
        const int t = i * N - (i * (i - 1)) / 2;
        c.diag_f[i] = LmController::clampd((c.ev.G1[t] + c.ev.G2[t]) * sci * sci, c.opt.min_lm_diagonal, c.opt.max_lm_diagonal);
    }
    __syncwarp();
    const double eps = 1.0 / (radius + 1.0);
    double a[N], invd[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const int jj = j <= i ? j : i;                           // row i only needs columns j <= i
        const int t = jj * N - (jj * (jj - 1)) / 2 + (i - jj);   // tri_index(N, jj, i)
        double sv = c.ev.G1[t] + c.ev.G2[t] * eps;
        if (exc) sv -= exc->S[t];
        a[j] = sv * (sci * c.scale_f[jj]);
        invd[j] = 0.0;
    }
    {
        const double dd = c.diag_f[i] / radius;                  // (sqrt(diag/radius))^2
#pragma unroll
        for (int j = 0; j < N; ++j) if (j == i) a[j] += dd;
    }
    double y = c.ev.h1[i] + c.ev.h2[i] * eps;
    if (exc) y -= exc->rhs[i];
    y *= sci;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double dk = __shfl_sync(0xffffffffu, a[k], k);      // pivot
        if (!(dk > 0.0)) ok = false;
        const double inv = fast_rsqrt(dk);
        invd[k] = inv;
        a[k] = (i == k) ? dk * inv : a[k] * inv;                  // l_kk = sqrt(d), l_ik = a_ik / l_kk
#pragma unroll
        for (int j = k + 1; j < N; ++j) {
            const double ljk = __shfl_sync(0xffffffffu, a[k], j);
            a[j] = fma(-a[k], ljk, a[j]);                         // only meaningful for i >= j
        }
    }
    // forward substitution L z = rhs
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double zk = __shfl_sync(0xffffffffu, y, k) * invd[k];
        if (i == k) y = zk; else if (i > k) y = fma(-a[k], zk, y);
    }
    // backward substitution L^T x = z: lane i needs column i of L
    if (lane < N) {
#pragma unroll
        for (int j = 0; j < N; ++j) if (j <= i) Lm[i][j] = a[j];
    }
    __syncwarp();
    double col[N];
#pragma unroll
    for (int k = 0; k < N; ++k) col[k] = (k > i) ? Lm[k][i] : 0.0;
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {
        const double xk = __shfl_sync(0xffffffffu, y, k) * invd[k];
        if (i == k) y = xk; else if (i < k) y = fma(-col[k], xk, y);
    }
    if (!isfinite(y)) ok = false;
    ok = __all_sync(0xffffffffu, ok);
    if (lane < N) c.delta_f[i] = -y * sci;                        // step = -y ; delta = step o scale
    __syncwarp();
    if (lane == 0) {
        c.reuse_diagonal = 1;
        nx = ok ? (int)LM_RUN_B : (int)c.invalid_step();
    }
    return __shfl_sync(0xffffffffu, nx, 0);
}

// ------------------------------------------------------------------------------------------
// The persistent kernel.  Phases: one INIT pass (evaluation at the start point), then one FUSED
// pass per LM iteration: the candidate step at x (back substitution, model cost change) and,
// speculatively, the complete evaluation at the candidate.  If the controller accepts the step the
// speculative sums ARE the next iteration's system; if it rejects, it re-solves from the stored
// radius-independent factors at a smaller radius -- either way the next phase is another FUSED pass.
// ------------------------------------------------------------------------------------------
constexpr unsigned long long kWatchdogNs = 4000000000ull;   // 4 s: a stuck grid barrier aborts the solve

// elected thread: queue the TMA bulk copies of one tile into a stage
__device__ __forceinline__ void issue_tile(const RefineData &D, const double *dx, int tile, Stage *st, uint64_t *bar)
{
    // the tile-blocked arrays and the depth buffers are padded to whole tiles: fixed copy sizes
    mbar_expect_tx(bar, (unsigned)(3 * kTile * sizeof(double2) + kTile * sizeof(double)));
    bulk_g2s(st->xy, D.blk + (size_t)tile * (3 * kTile), (unsigned)(3 * kTile * sizeof(double2)), bar);
    bulk_g2s(st->d, dx + (size_t)tile * kTile, (unsigned)(kTile * sizeof(double)), bar);
}

// Deterministic summation order for the listed pixels: bitonic sort (shared memory, whole CTA) of
// keys[k] = (residual-block index << 32) | list slot.  false: the list does not fit (more than `cap`
// listed pixels) and is summed in slot order -- correct, but then not bit-reproducible run to run.
__device__ __forceinline__ bool sort_exceptions(const ExcEntry *list, int ne, unsigned long long *keys, int cap, int tid)
{
    int npad = 1;
    while (npad < ne) npad <<= 1;
    if (npad > cap) return false;
    for (int k = tid; k < npad; k += kThreads)
        keys[k] = (k < ne) ? ((unsigned long long)(unsigned int)(int)__ldcg(&list[k].key) << 32) | (unsigned int)k : ~0ull;
    __syncthreads();
    for (int size = 2; size <= npad; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (npad >> 1); i += kThreads) {
                const int lo = 2 * stride * (i / stride) + (i % stride), hi = lo + stride;
                const unsigned long long a = keys[lo], b = keys[hi];
                if ((a > b) == ((lo & size) == 0)) { keys[lo] = b; keys[hi] = a; }
            }
            __syncthreads();
        }
    return true;
}

template <int NF>
__global__ void __launch_bounds__(kThreads, 1)
k_lm_persistent(RefineData D, double *d0, double *d1, LmShared *sh, double *partials, ExcEntry *exc, unsigned int exc_cap,
                const double *z_in, int z_stride, double *out, int invert_out, double *zstats)
{
    using A = Acc<NF>;
    constexpr int LD = kRowVals<NF>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Stage *stages = reinterpret_cast<Stage *>(smem_raw);
    __shared__ __align__(8) uint64_t full[kStages];
    __shared__ PhaseParams P;
    __shared__ LmController s_ctl;
    __shared__ double fin[LD];
    __shared__ double part[kWarps][LD];
    __shared__ ExcSums s_exc;
    __shared__ double s_L[7][8], s_y[8];                          // controller solve scratch
    __shared__ int s_flag[6];                                     // [0] is_last, [1] next, [2] n_exc, [3] accepted, [4] cur_list

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    const int NT = (D.m + kTile - 1) / kTile;
    const int n_my = ((int)blockIdx.x < NT) ? (NT - 1 - (int)blockIdx.x) / G + 1 : 0;
    unsigned int gen = 0;
    unsigned int consumed = 0;                                    // tiles consumed by this CTA since kernel start

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < (int)(sizeof(Motion) / sizeof(int)))
        reinterpret_cast<int *>(&P.base)[tid] = __ldcg(reinterpret_cast<const int *>(&sh->base) + tid);
    __syncthreads();

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
        const bool run_init = (P.next == LM_RUN_A);
        double *dx = P.which_x ? d1 : d0;
        double *dcand = P.which_x ? d0 : d1;
        double *row = partials + (size_t)blockIdx.x * A::NV;
        // exceptions of the evaluation point go to the current list (INIT) or to the speculative one (FUSED)
        const int elist = run_init ? P.cur_list : (P.cur_list ^ 1);
        unsigned int *n_exc = &sh->n_exc[elist];
        ExcEntry *elist_p = exc + (size_t)elist * exc_cap;
        const unsigned long long t_begin = (blockIdx.x == 0 && tid == 0) ? globaltimer() : 0ull;

        // ---- prologue: fill the ring
        if (tid == 0) {
            fence_proxy_async();
            const int pre = n_my < kStages ? n_my : kStages;
            for (int k = 0; k < pre; ++k) {
                const unsigned g = consumed + (unsigned)k;
                issue_tile(D, dx, (int)blockIdx.x + k * G, &stages[g % kStages], &full[g % kStages]);
            }
        }

        double acc[TAcc<NF>::NV];
#pragma unroll
        for (int j = 0; j < TAcc<NF>::NV; ++j) acc[j] = 0.0;
        const double c2 = 2.0 / (2.0 + P.mot.k), c2c = 2.0 / (2.0 + P.cand.k);
        const double rfac = P.radius / (P.radius + 1.0), inv_radius = 1.0 / P.radius;

        // one 512-block tile per iteration: every thread works on two residual blocks at a time
        for (int k = 0; k < n_my; ++k) {
            const unsigned g = consumed + (unsigned)k;
            const int s = (int)(g % kStages);
            const int base_i = ((int)blockIdx.x + k * G) * kTile;
            Loaded L[2];
            bool valid[2];
            int idx[2];
            mbar_wait(&full[s], (g / kStages) & 1u);
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int j = tid + p * kThreads;
                idx[p] = base_i + j;
                valid[p] = idx[p] < D.m;
                L[p].p = stages[s].xy[j]; L[p].u = stages[s].uu[j]; L[p].a = stages[s].aa[j]; L[p].d = stages[s].d[j];
                if (!valid[p]) { L[p].p = make_double2(0.0, 0.0); L[p].u = L[p].p; L[p].a = make_double2(1.0, 0.0); L[p].d = 1.0; }
            }
            // stage s may be refilled once every thread has read it: the consumer warps only ARRIVE on
            // the stage's named barrier and run on; warp 0 (the producer) waits for the 256 arrivals
            if (warp == 0) asm volatile("bar.sync %0, %1;" ::"r"(1 + s), "r"(kThreads) : "memory");
            else asm volatile("bar.arrive %0, %1;" ::"r"(1 + s), "r"(kThreads) : "memory");
            if (tid == 0 && k + kStages < n_my) {
                fence_proxy_async();
                issue_tile(D, dx, (int)blockIdx.x + (k + kStages) * G, &stages[s], &full[s]);
            }
            if (run_init) {
                const double dv[2] = {L[0].d, L[1].d};
                eval_pair<NF>(L, dv, valid, base_i, P.mot, c2, P.first != 0, P, acc, n_exc, &sh->exc_overflow, elist_p, exc_cap);
            } else {
                double dc[2];
                step_pair<NF>(L, valid, idx, P, c2, rfac, inv_radius, acc, dcand, dc);
                eval_pair<NF>(L, dc, valid, base_i, P.cand, c2c, false, P, acc, n_exc, &sh->exc_overflow, elist_p, exc_cap);
            }
        }
        consumed += (unsigned)n_my;
        __syncthreads();
        const unsigned long long t_loop = t_begin ? globaltimer() : 0ull;
        cta_reduce_roles<NF, LD>(acc, part, row);
        if (t_begin) {
            const unsigned long long t2 = globaltimer();
            atomicAdd(&sh->t_phase[run_init ? 4 : 7], t_loop - t_begin); atomicAdd(&sh->t_phase[run_init ? 5 : 8], t2 - t_loop);
        }

        // ---- grid barrier: the last CTA to arrive reduces the rows and runs the controller
        if (tid == 0) {
            __threadfence();
            const unsigned int ticket = atomicAdd(&sh->arrive, 1u);
            s_flag[0] = (ticket == (gen + 1u) * (unsigned)G - 1u) ? 1 : 0;
        }
        __syncthreads();
        if (s_flag[0]) {
            __threadfence();
            const unsigned long long t_ctl = (tid == 0) ? globaltimer() : 0ull;
            constexpr int nv = A::NV, ns = A::NS;
            // final reduce: warp w sums rows w, w+8, ...; lanes cover the columns (coalesced).  All loads of a
            // warp (up to 19 rows x 3 column groups) and the controller state are issued before anything is
            // combined -- one L2 round trip instead of one per batch; the sums run in a fixed order.
            // (the exception-list counters too: thread 0 needs one of them in the middle of the controller logic)
            const unsigned int ne_pre0 = (tid == 0) ? __ldcg(&sh->n_exc[0]) : 0u, ne_pre1 = (tid == 0) ? __ldcg(&sh->n_exc[1]) : 0u;
            int ctl_w[(sizeof(LmController) / sizeof(int) + kThreads - 1) / kThreads];
#pragma unroll
            for (int q = 0; q < (int)(sizeof(ctl_w) / sizeof(int)); ++q) {
                const int w = tid + q * kThreads;
                ctl_w[q] = (w < (int)(sizeof(LmController) / sizeof(int))) ? __ldcg(reinterpret_cast<const int *>(&sh->ctl) + w) : 0;
            }
            {
                constexpr int kRowsPerWarp = (kNumSMsB200 + kWarps - 1) / kWarps;      // 19
                double v[3] = {0.0, 0.0, 0.0};
                if (G <= kNumSMsB200) {
                    double t[kRowsPerWarp][3];
#pragma unroll
                    for (int u = 0; u < kRowsPerWarp; ++u) {
                        const int b = u * kWarps + warp;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const int j = lane + 32 * c;
                            t[u][c] = (b < G && j < nv) ? __ldcg(partials + (size_t)b * A::NV + j) : 0.0;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kRowsPerWarp; ++u)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const int j = lane + 32 * c;
                            v[c] = (j < ns) ? v[c] + t[u][c] : fmax(v[c], t[u][c]);
                        }
                } else {
                    for (int b = warp; b < G; b += kWarps)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const int j = lane + 32 * c;
                            const double x = (j < nv) ? __ldcg(partials + (size_t)b * A::NV + j) : 0.0;
                            v[c] = (j < ns) ? v[c] + x : fmax(v[c], x);
                        }
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) { const int j = lane + 32 * c; if (j < nv) part[warp][j] = v[c]; }
            }
            // controller state: global -> shared
#pragma unroll
            for (int q = 0; q < (int)(sizeof(ctl_w) / sizeof(int)); ++q) {
                const int w = tid + q * kThreads;
                if (w < (int)(sizeof(LmController) / sizeof(int))) reinterpret_cast<int *>(&s_ctl)[w] = ctl_w[q];
            }
            __syncthreads();
            if (tid < nv) {
                double x = part[0][tid];
                if (tid < ns) for (int w = 1; w < kWarps; ++w) x += part[w][tid];
                else          for (int w = 1; w < kWarps; ++w) x = fmax(x, part[w][tid]);
                fin[tid] = x;
            }
            __syncthreads();
            const unsigned long long t_fin = (tid == 0) ? globaltimer() : 0ull;
            // ---- FUSED: judge the candidate first
            if (tid == 0) {
                int accepted = run_init ? 1 : 0;
                LmNext nx = LM_RUN_A;
                if (!run_init) {
                    CandSums c;
                    c.mcc = fin[A::oMCC]; c.step_sq = fin[A::oSTEP]; c.cand_cost = 0.5 * fin[0];
                    c.bad_step = fin[A::iBADSTEP]; c.bad_cand = fin[A::iBADRES];
                    nx = s_ctl.on_candidate(c);
                    accepted = (nx == LM_RUN_A) ? 1 : 0;
                }
                s_flag[1] = (int)nx;
                s_flag[3] = accepted;
            }
            __syncthreads();
            if (s_flag[3]) {
                // the evaluation sums of this pass describe the (new) current point: EvalSums in place
                if (tid < kTri) {
                    s_ctl.ev.G1[tid] = (tid < A::TRI) ? fin[A::oG1 + (tid < A::TRI ? tid : 0)] : 0.0;
                    s_ctl.ev.G2[tid] = (tid < A::TRI) ? fin[A::oG2 + (tid < A::TRI ? tid : 0)] : 0.0;
                }
                if (tid < kMaxNF) {
                    s_ctl.ev.h1[tid] = (tid < NF) ? fin[A::oH1 + (tid < NF ? tid : 0)] : 0.0;
                    s_ctl.ev.h2[tid] = (tid < NF) ? fin[A::oH2 + (tid < NF ? tid : 0)] : 0.0;
                }
                if (tid == 32) {
                    s_ctl.ev.cost = 0.5 * fin[0]; s_ctl.ev.sumsq_d = fin[1]; s_ctl.ev.gmax_e = fin[A::iGMAX];
                    s_ctl.ev.bad = fin[A::iBAD]; s_ctl.ev.ee_max = fin[A::iEEMAX];
                }
                __syncthreads();
            }
            if (tid == 0) {
                // exception lists: on acceptance the speculative list becomes the current one
                const int cur = run_init ? P.cur_list : (s_flag[3] ? (P.cur_list ^ 1) : P.cur_list);
                s_flag[4] = cur;
                const unsigned int ne = cur ? ne_pre1 : ne_pre0;
                s_flag[2] = (int)(ne < exc_cap ? ne : exc_cap);
                sh->n_exc[cur ^ 1] = 0u;                       // the other list is rebuilt by the next pass
            }
            __syncthreads();
            // the pixel ring is idle during the controller section: its shared memory holds the sort keys
            unsigned long long *xkeys = reinterpret_cast<unsigned long long *>(smem_raw);
            bool xsorted = false;
            if constexpr (NF > 0) if (s_flag[2] > 0)
                xsorted = sort_exceptions(exc + (size_t)s_flag[4] * exc_cap, s_flag[2], xkeys, 16384, tid);
            if (s_flag[3]) {
                // listed pixels: their radius-independent part F^TF, F^Tr joins G1, h1 of the new point
                if constexpr (NF > 0) if (s_flag[2] > 0) {
                    const ExcEntry *cur_exc = exc + (size_t)s_flag[4] * exc_cap;
                    double a[kExcVals];
#pragma unroll
                    for (int j = 0; j < kExcVals; ++j) a[j] = 0.0;
                    for (int k = tid; k < s_flag[2]; k += kThreads) {
                        const int slot = xsorted ? (int)(unsigned int)(xkeys[k] & 0xffffffffull) : k;
                        ExcEntry X;
                        for (int w = 0; w < (int)(sizeof(ExcEntry) / sizeof(double)); ++w)
                            reinterpret_cast<double *>(&X)[w] = __ldcg(reinterpret_cast<const double *>(cur_exc + slot) + w);
                        int t = 0;
#pragma unroll
                        for (int j = 0; j < NF; ++j) {
                            a[kTri + j] += fma(X.F0[j], X.r0, X.F1[j] * X.r1);
#pragma unroll
                            for (int c = j; c < NF; ++c, ++t) a[t] += fma(X.F0[j], X.F0[c], X.F1[j] * X.F1[c]);
                        }
                    }
                    cta_reduce_sums<kExcVals, LD>(a, part, fin);
                    if (tid < A::TRI) s_ctl.ev.G1[tid] += fin[tid];
                    if (tid < NF) s_ctl.ev.h1[tid] += fin[kTri + tid];
                    __syncthreads();
                }
                if (warp == 0) {
                    const int nx = ctl_on_eval<NF>(s_ctl);
                    if (lane == 0) s_flag[1] = nx;
                }
                __syncthreads();
            }
            // ---- (re)solve at the current radius; the clamped-pixel correction is summed by the whole CTA
            while (s_flag[1] == (int)LM_SOLVE) {
                const int ne = s_flag[2];
                if constexpr (NF > 0) if (ne > 0) {
                    const ExcEntry *cur_exc = exc + (size_t)s_flag[4] * exc_cap;
                    const double R = s_ctl.radius, lo = s_ctl.opt.min_lm_diagonal, hi = s_ctl.opt.max_lm_diagonal;
                    double a[kExcVals];
#pragma unroll
                    for (int j = 0; j < kExcVals; ++j) a[j] = 0.0;
                    for (int k = tid; k < ne; k += kThreads) {
                        const int slot = xsorted ? (int)(unsigned int)(xkeys[k] & 0xffffffffull) : k;
                        ExcEntry X;
                        for (int w = 0; w < (int)(sizeof(ExcEntry) / sizeof(double)); ++w)
                            reinterpret_cast<double *>(&X)[w] = __ldcg(reinterpret_cast<const double *>(cur_exc + slot) + w);
                        const double q = X.se2 / (X.ees + fmin(fmax(X.ees, lo), hi) / R);
                        const double er = fma(X.e0, X.r0, X.e1 * X.r1);
                        double fe[NF > 0 ? NF : 1];
#pragma unroll
                        for (int j = 0; j < NF; ++j) fe[j] = fma(X.F0[j], X.e0, X.F1[j] * X.e1);
                        int t = 0;
#pragma unroll
                        for (int j = 0; j < NF; ++j) {
                            const double qf = q * fe[j];
                            a[kTri + j] = fma(qf, er, a[kTri + j]);
#pragma unroll
                            for (int c = j; c < NF; ++c, ++t) a[t] = fma(qf, fe[c], a[t]);
                        }
                    }
                    cta_reduce_sums<kExcVals, LD>(a, part, fin);
                    if (tid < kTri) s_exc.S[tid] = fin[tid];
                    if (tid < kMaxNF) s_exc.rhs[tid] = fin[kTri + tid];
                    __syncthreads();
                }
                if (warp == 0) {
                    const int nx = ctl_solve<NF>(s_ctl, (NF > 0 && ne > 0) ? &s_exc : nullptr, s_L, s_y);
                    if (lane == 0) s_flag[1] = nx;
                }
                __syncthreads();
            }
            // ---- publish the next phase
            if (tid == 0) {
                const unsigned long long t_solved = globaltimer();
                atomicAdd(&sh->t_phase[run_init ? 10 : 11], t_solved - t_fin);   // controller logic only (after the row reduction)
                const LmNext nx = (LmNext)s_flag[1];                     // LM_RUN_B (another fused pass) or LM_DONE
                if (!run_init && s_flag[3]) sh->bc.which_x = P.which_x ^ 1;   // the candidate became x
                Motion mo = P.base, ca = P.base;
                if (NF >= 6) for (int j = 0; j < 3; ++j) {
                    mo.v[j] = s_ctl.f[j]; mo.w[j] = s_ctl.f[3 + j];
                    ca.v[j] = s_ctl.f[j] + s_ctl.delta_f[j]; ca.w[j] = s_ctl.f[3 + j] + s_ctl.delta_f[3 + j];
                }
                if (NF == 7) { mo.k = s_ctl.f[6]; ca.k = s_ctl.f[6] + s_ctl.delta_f[6]; }
                if (nx == LM_DONE && s_ctl.termination == RSDSFM_FAILURE) mo = P.base;   // Ceres restores the start values
                sh->bc.mot = mo; sh->bc.cand = ca;
                for (int j = 0; j < kMaxNF; ++j) sh->bc.delta_f[j] = s_ctl.delta_f[j];
                sh->bc.radius = s_ctl.radius;
                sh->bc.ee_fast_min = s_ctl.ee_fast_min;
                sh->bc.first = 0;
                sh->bc.cur_list = s_flag[4];
                sh->bc.next = (int)nx;
                if (t_begin) {
                    const unsigned long long dt = globaltimer() - t_begin;
                    atomicAdd(&sh->t_phase[run_init ? 0 : 2], dt); atomicAdd(&sh->t_phase[run_init ? 1 : 3], 1ull);
                }
            }
            for (int w = tid; w < (int)(sizeof(LmController) / sizeof(int)); w += kThreads)
                reinterpret_cast<int *>(&sh->ctl)[w] = reinterpret_cast<const int *>(&s_ctl)[w];
            __syncthreads();
            if (tid == 0) {
                atomicAdd(&sh->t_phase[run_init ? 6 : 9], globaltimer() - t_ctl);
                __threadfence();
                st_release(&sh->generation, gen + 1u);
            }
        }
        // ---- everybody waits for the controller's release
        if (tid == 0) {
            const unsigned long long t0 = globaltimer();
            while (ld_acquire(&sh->generation) <= gen) {
                __nanosleep(32);
                if (globaltimer() - t0 > kWatchdogNs) { sh->error = 1; break; }
            }
            if (t_begin && !s_flag[0]) {
                const unsigned long long dt = globaltimer() - t_begin;
                atomicAdd(&sh->t_phase[run_init ? 0 : 2], dt); atomicAdd(&sh->t_phase[run_init ? 1 : 3], 1ull);   // fire and forget
            }
        }
        __syncthreads();
        gen++;
    }

    // ---- epilogue: write the result (z = 1/d for a9, d for a8).  On FAILURE Ceres restores the
    // start values (solver.cc Minimize): 1/z_in for a9 (double reciprocal, :213/:247), 1.0 for a8.
    const bool failed = (__ldcg(&sh->ctl.termination) == RSDSFM_FAILURE) || P.error;
    const double *dfin = P.which_x ? d1 : d0;
    // zstats (nullable): per-CTA rows {sum z, max z, max -z} of what was written, for the sign fix
    // and depth range of main.cc:466-489 -- saves the rectification stage a pass over z
    double zs[1] = {0.0}, zm[2] = {-INFINITY, -INFINITY};
    for (int i = blockIdx.x * kThreads + tid; i < D.m; i += G * kThreads) {
        double dv;
        if (failed) dv = z_in ? 1.0 / z_in[(size_t)i * z_stride] : 1.0;
        else dv = dfin[i];
        const double o = invert_out ? 1.0 / dv : dv;
        out[i] = o;
        zs[0] += o; zm[0] = fmax(zm[0], o); zm[1] = fmax(zm[1], -o);
    }
    if (zstats) block_reduce_store<1, 2>(zs, zm, zstats);
}

// ==========================================================================================
// k_lm_solve -- second-generation persistent solver (replaces k_lm_persistent above, which is
// kept for A/B measurements: RSDSFM_LM_VARIANT=1).
//
//  * REPLICATED CONTROLLER.  Every CTA publishes its row of partial sums, arrives ONCE on a grid
//    counter, then reads all rows itself, sums them in the same fixed order and runs the same
//    Ceres logic on its own shared-memory copy of the controller state.  All CTAs take bit-identical
//    decisions; there is no "last CTA", no serial publish/release step and no control block
//    round trip through global memory between phases.  (It is also what a multi-GPU row split
//    needs: peers only have to make their rows visible.)
//  * PREFETCH ACROSS THE BARRIER.  Right after arriving, a CTA queues the TMA loads of the next
//    phase's first tiles: the 24 KB static part unconditionally, the 4 KB inverse-depth part from the
//    buffer the next phase reads IF THE STEP IS ACCEPTED (this CTA wrote those depths itself).  A
//    rejected step drains the ring and reloads (rare).
//  * EMPTY/FULL MBARRIER RING.  Consumers release a stage by arriving on its `empty` mbarrier (one
//    arrival per warp); the producer thread refills one tile behind the consumers with a
//    non-blocking test, so no warp ever waits at a CTA-wide barrier inside the sweep.
//  * ONE PIXEL PER THREAD PER STEP, SOFTWARE-PIPELINED.  The rank-1 Schur updates of step i-1 (70
//    independent FMAs on vectors held in registers) are issued together with the latency-bound
//    residual / Jacobian chain of step i, which is what keeps the FP64 pipe fed with two warps per
//    sub-partition.  Fewer FP64 instructions per residual block, too: the candidate's A v and B w
//    are updated from the step's increments, the two projected Jacobian vectors share their
//    products, flags and maxima are tracked with integer instructions.
// ==========================================================================================
constexpr int kExcSlots = 3;

template <int NF>
struct Held {                                   // projected, 1/|e|-scaled Jacobian vectors of the previous step
    double kv[NF > 0 ? NF : 1], pv[NF > 0 ? NF : 1], ks, ps;
};

struct SweepScalars {                           // per-thread non-FP64 accumulators of a sweep
    unsigned long long gmax, eemax;             // bit patterns of max |e^T r|, max e^Te (non-negative doubles order like integers)
    unsigned flags;                             // 1: residual not finite, 2: residual or Jacobian not finite, 4: depth step not finite
};

__device__ __forceinline__ bool not_finite(double x) { return (__double2hiint(x) & 0x7ff00000) == 0x7ff00000; }
__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

__device__ __forceinline__ bool mbar_test(uint64_t *bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Evaluation of one residual block at the point described by (ak, beta, a, b, d): residual, depth
// column, cost / gradient sums, and the two projected Jacobian vectors of the lane-pair scheme:
// H.kv / H.ks = this lane's direction (even lanes n, odd lanes e), sv / ss = the direction the partner keeps.
template <int NF>
__device__ __forceinline__ void eval_pixel(const Loaded &L, bool valid, int index, double xy, double xx1, double yy1, double ak,
                                           double beta, double c2, double a0, double a1, double b0, double b1, double d_in,
                                           bool first, const PhaseParams &P, const Motion &mot, double (&acc)[TAcc<NF>::NV],
                                           SweepScalars &S, double (&kv)[NF > 0 ? NF : 1], double (&sv)[NF > 0 ? NF : 1], double &ks,
                                           double &ss, unsigned int *n_exc, unsigned int *overflow, ExcEntry *exc, unsigned int exc_cap)
{
    const double x = L.p.x, y = L.p.y;
    double d = d_in;
    const double p0 = fma(d, a0, b0), p1 = fma(d, a1, b1);
    double r0 = fma(-beta, p0, L.u.x), r1 = fma(-beta, p1, L.u.y);
    double e0 = -beta * a0, e1 = -beta * a1;
    double ee = fma(e0, e0, e1 * e1);
    if (!valid) { r0 = 0.0; r1 = 0.0; e0 = 0.0; e1 = 0.0; ee = 0.0; d = 0.0; }
    acc[0] = fma(r0, r0, fma(r1, r1, acc[0]));
    acc[1] = fma(d, d, acc[1]);
    const double re = fma(e0, r0, e1 * r1);                           // e^T r
    S.gmax = umax64(S.gmax, (unsigned long long)__double_as_longlong(fabs(re)));
    S.eemax = umax64(S.eemax, (unsigned long long)__double_as_longlong(ee));
    if (not_finite(r0 + r1)) S.flags |= 3u;
    if (not_finite(ee)) S.flags |= 2u;
    // is the LM diagonal of this depth certainly not clamped?  (first evaluation: the Jacobi scale is
    // 1/(1+|e|) of this very point; later: global lower bound of the scales)
    bool fast;
    if (first) {
        const double se = 1.0 / (1.0 + sqrt(ee));
        const double ees = ee * se * se;
        fast = (ees >= P.min_diag && ees <= P.max_diag);
    } else {
        fast = (ee >= P.ee_fast_min && ee <= P.max_diag);
    }
    ks = 0.0; ss = 0.0;
    if (NF > 0) {
        const double mu = (valid && fast) ? fast_rsqrt(ee) : 0.0;       // 1/|e|
        const double c = mu * e0, s = mu * e1;                          // unit depth-column direction
        const bool e_role = (threadIdx.x & 1) != 0;
        const double mc = e_role ? c : -s, ms = e_role ? s : c;         // the direction this lane keeps: e or n = (-s, c)
        // F^T (mc, ms) and F^T (-ms, mc) share their products (F = -beta [d A | B | (dbeta/beta) p])
        const double P0 = beta * mc, P1 = beta * ms;
        const double dP0 = d * P0, dP1 = d * P1;
        const double t1 = fma(x, P0, y * P1), t2 = fma(y, P0, -x * P1);
        kv[0] = -dP0;            sv[0] = dP1;
        kv[1] = -dP1;            sv[1] = -dP0;
        kv[2] = d * t1;          sv[2] = d * t2;
        kv[3] = fma(y, t1, P1);  sv[3] = fma(y, t2, P0);
        kv[4] = -fma(x, t1, P0); sv[4] = fma(-x, t2, P1);
        kv[5] = t2;              sv[5] = -t1;
        if (NF == 7) {
            const double dbeta = c2 * fma(-ak, 0.5 * c2, L.a.y);
            kv[6] = -dbeta * fma(p0, mc, p1 * ms);
            sv[6] = -dbeta * fma(p1, mc, -p0 * ms);
        }
        ks = fma(mc, r0, ms * r1);
        ss = fma(mc, r1, -ms * r0);
        if (valid && !fast) {
            // rare: the pixel is listed and handled exactly by the controller (see eval_slow); its vectors are 0 here.
            // The out-of-line call takes the pixel by address: hand it a copy made HERE.
            const Loaded Lc = L;
            eval_slow<NF>(Lc, d_in, index, mot, c2, first, P.base, n_exc, overflow, exc, exc_cap);
        }
    }
}

// INIT phase: evaluation at the start point.
template <int NF>
__device__ __forceinline__ void init_pixel(const Loaded &L, bool valid, int index, const PhaseParams &P, double c2,
                                           double (&acc)[TAcc<NF>::NV], SweepScalars &S, double (&kv)[NF > 0 ? NF : 1],
                                           double (&sv)[NF > 0 ? NF : 1], double &ks, double &ss, unsigned int *n_exc,
                                           unsigned int *overflow, ExcEntry *exc, unsigned int exc_cap)
{
    const double x = L.p.x, y = L.p.y;
    const double xy = x * y, xx1 = fma(x, x, 1.0), yy1 = fma(y, y, 1.0);
    const double ak = fma(P.mot.k, L.a.y, L.a.x), beta = c2 * ak;
    const double a0 = fma(-x, P.mot.v[2], P.mot.v[0]), a1 = fma(-y, P.mot.v[2], P.mot.v[1]);
    const double b0 = fma(-xy, P.mot.w[0], fma(xx1, P.mot.w[1], -y * P.mot.w[2]));
    const double b1 = fma(-yy1, P.mot.w[0], fma(xy, P.mot.w[1], x * P.mot.w[2]));
    eval_pixel<NF>(L, valid, index, xy, xx1, yy1, ak, beta, c2, a0, a1, b0, b1, L.d, P.first != 0, P, P.mot, acc, S, kv, sv, ks, ss,
                   n_exc, overflow, exc, exc_cap);
}

// FUSED phase: candidate step at x (depth back-substitution delta_d = -q e^T (r + F delta_f), model cost
// change, |step|^2, candidate depth written to d_cand) followed by the evaluation at the candidate.
template <int NF>
__device__ __forceinline__ void fused_pixel(const Loaded &L, bool valid, int index, const PhaseParams &P, double c2, double c2c,
                                            double rfac, double inv_radius, double (&acc)[TAcc<NF>::NV], SweepScalars &S,
                                            double *__restrict__ d_cand, double (&kv)[NF > 0 ? NF : 1],
                                            double (&sv)[NF > 0 ? NF : 1], double &ks, double &ss, unsigned int *n_exc,
                                            unsigned int *overflow, ExcEntry *exc, unsigned int exc_cap)
{
    using T = TAcc<NF>;
    const double x = L.p.x, y = L.p.y, d = L.d;
    const double xy = x * y, xx1 = fma(x, x, 1.0), yy1 = fma(y, y, 1.0);
    // ---- at x
    const double ak = fma(P.mot.k, L.a.y, L.a.x), beta = c2 * ak;
    const double a0 = fma(-x, P.mot.v[2], P.mot.v[0]), a1 = fma(-y, P.mot.v[2], P.mot.v[1]);
    const double b0 = fma(-xy, P.mot.w[0], fma(xx1, P.mot.w[1], -y * P.mot.w[2]));
    const double b1 = fma(-yy1, P.mot.w[0], fma(xy, P.mot.w[1], x * P.mot.w[2]));
    const double p0 = fma(d, a0, b0), p1 = fma(d, a1, b1);
    const double r0 = fma(-beta, p0, L.u.x), r1 = fma(-beta, p1, L.u.y);
    const double e0 = -beta * a0, e1 = -beta * a1;
    const double ee = fma(e0, e0, e1 * e1);
    // q = s_e^2 / (s_e^2 e^Te + clamp(s_e^2 e^Te)/radius)  ( = radius/((radius+1) e^Te) when not clamped )
    double q = fast_rcp(ee) * rfac;
    if (!(ee >= P.ee_fast_min && ee <= P.max_diag)) {
        const double se = depth_scale_at_start(L, P.base);
        const double se2 = se * se, ees = ee * se2;
        q = se2 / (ees + fmin(fmax(ees, P.min_diag), P.max_diag) * inv_radius);
    }
    // F delta_f = -beta (d A dv + B dw) - dbeta p dk; the increments of A v and B w are reused for the candidate
    double da0 = 0.0, da1 = 0.0, db0 = 0.0, db1 = 0.0, m0 = 0.0, m1 = 0.0;
    if (NF >= 6) {
        const double *df = P.delta_f;
        da0 = fma(-x, df[2], df[0]); da1 = fma(-y, df[2], df[1]);
        db0 = fma(-xy, df[3], fma(xx1, df[4], -y * df[5]));
        db1 = fma(-yy1, df[3], fma(xy, df[4], x * df[5]));
        m0 = -beta * fma(d, da0, db0);
        m1 = -beta * fma(d, da1, db1);
        if (NF == 7) {
            const double dbk = c2 * fma(-ak, 0.5 * c2, L.a.y) * df[6];
            m0 = fma(-dbk, p0, m0);
            m1 = fma(-dbk, p1, m1);
        }
    }
    const double delta_e = -q * fma(e0, r0 + m0, e1 * (r1 + m1));
    const double j0 = fma(e0, delta_e, m0), j1 = fma(e1, delta_e, m1);                    // J delta
    double dc = d + delta_e;
    const double dd = d - dc;
    if (valid) {
        d_cand[index] = dc;
        acc[T::oMCC] += fma(j0, fma(0.5, j0, r0), j1 * fma(0.5, j1, r1));
        acc[T::oSTEP] = fma(dd, dd, acc[T::oSTEP]);
        if (not_finite(delta_e)) S.flags |= 4u;
    } else {
        dc = 1.0;
    }
    // ---- at the candidate
    const double akc = fma(P.cand.k, L.a.y, L.a.x), betac = c2c * akc;
    eval_pixel<NF>(L, valid, index, xy, xx1, yy1, akc, betac, c2c, a0 + da0, a1 + da1, b0 + db0, b1 + db1, dc, false, P, P.cand, acc, S,
                   kv, sv, ks, ss, n_exc, overflow, exc, exc_cap);
}

template <int NF>
__device__ __forceinline__ void accumulate_held(const Held<NF> &h, double (&acc)[TAcc<NF>::NV])
{
    using T = TAcc<NF>;
    int t = 0;
#pragma unroll
    for (int j = 0; j < NF; ++j) {
        acc[T::oH + j] = fma(h.kv[j], h.ks, fma(h.pv[j], h.ps, acc[T::oH + j]));
#pragma unroll
        for (int c = j; c < NF; ++c, ++t) acc[T::oK + t] = fma(h.kv[j], h.kv[c], fma(h.pv[j], h.pv[c], acc[T::oK + t]));
    }
}

struct SolveArgs {
    RefineData D;
    double *d0, *d1;
    LmShared *sh;
    double *partials;            // [2][gridDim.x][Acc<NF>::NV]: rows of even / odd phases
    ExcEntry *exc;               // [kExcSlots][exc_cap]
    unsigned int exc_cap;
    const double *z_in;
    int z_stride;
    double *out;
    int invert_out;
    double *zstats;
};

template <int NF>
__global__ void __launch_bounds__(kThreads, 1) k_lm_solve(const SolveArgs A_)
{
    using A = Acc<NF>;
    using T = TAcc<NF>;
    constexpr int LD = kRowVals<NF>;
    constexpr int NFa = NF > 0 ? NF : 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Stage *stages = reinterpret_cast<Stage *>(smem_raw);
    __shared__ __align__(8) uint64_t full[kStages], empty[kStages];
    __shared__ PhaseParams P;
    __shared__ LmController s_ctl;
    __shared__ double fin[LD];
    __shared__ double part[kWarps][LD];
    __shared__ ExcSums s_exc;
    __shared__ double s_L[7][8], s_y[8];
    __shared__ int s_flag[8];     // [1] next, [2] n_exc of the current list, [3] accepted, [4] current slot, [5] error
    __shared__ unsigned int s_ne[kExcSlots];

    const RefineData D = A_.D;
    double *const d0 = A_.d0, *const d1 = A_.d1;
    LmShared *const sh = A_.sh;
    ExcEntry *const exc = A_.exc;
    const unsigned int exc_cap = A_.exc_cap;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    const int NT = (D.m + kTile - 1) / kTile;
    const int n_my = ((int)blockIdx.x < NT) ? (NT - 1 - (int)blockIdx.x) / G + 1 : 0;
    const int pre = n_my < kStages ? n_my : kStages;              // tiles queued ahead of a phase
    unsigned int gen = 0;
    unsigned int consumed = 0;                                    // tile uses consumed by this CTA since kernel start
    unsigned int issued = 0;                                      // (thread 0) tile uses queued since kernel start
    // exception lists: cur = list of the current point, spec = list the FUSED evaluation appends to,
    // zero = list that thread 0 of CTA 0 clears during this phase (it becomes `spec` of the next phase)
    int slot_cur = 0, slot_spec = 1, slot_zero = 2;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // replicated controller state + first phase parameters (written by k_lm_begin)
    for (int w = tid; w < (int)(sizeof(LmController) / sizeof(int)); w += kThreads)
        reinterpret_cast<int *>(&s_ctl)[w] = __ldcg(reinterpret_cast<const int *>(&sh->ctl) + w);
    if (tid < (int)(sizeof(Bcast) / sizeof(int)))
        reinterpret_cast<int *>(&P)[tid] = __ldcg(reinterpret_cast<const int *>(&sh->bc) + tid);
    if (tid < (int)(sizeof(Motion) / sizeof(int)))
        reinterpret_cast<int *>(&P.base)[tid] = __ldcg(reinterpret_cast<const int *>(&sh->base) + tid);
    if (tid == 64) {
        P.min_diag = __ldcg(&sh->ctl.opt.min_lm_diagonal); P.max_diag = __ldcg(&sh->ctl.opt.max_lm_diagonal);
        P.error = 0;
    }
    __syncthreads();

    // queue tile uses [issued, upto): use u is tile (u - phase_base) of the phase that reads depth buffer `dsrc`
    auto queue_uses = [&](unsigned int upto, unsigned int phase_base, const double *dsrc, bool blocking) {
        while (issued < upto) {
            const int s = (int)(issued % kStages);
            if (issued >= (unsigned)kStages) {
                const unsigned par = ((issued / kStages) - 1u) & 1u;           // completion of the stage's previous use
                if (blocking) mbar_wait(&empty[s], par);
                else if (!mbar_test(&empty[s], par)) break;
            }
            fence_proxy_async();
            issue_tile(D, dsrc, (int)blockIdx.x + (int)(issued - phase_base) * G, &stages[s], &full[s]);
            ++issued;
        }
    };
    if (tid == 0 && P.next != LM_DONE) queue_uses((unsigned)pre, 0u, P.which_x ? d1 : d0, true);

    for (;;) {
        if (P.next == LM_DONE || P.error) break;
        const bool run_init = (P.next == LM_RUN_A);
        double *dx = P.which_x ? d1 : d0;
        double *dcand = P.which_x ? d0 : d1;
        const int elist = run_init ? slot_cur : slot_spec;
        unsigned int *n_exc = &sh->n_exc[elist];
        ExcEntry *elist_p = exc + (size_t)elist * exc_cap;
        const unsigned long long t_begin = (blockIdx.x == 0 && tid == 0) ? globaltimer() : 0ull;
        if (blockIdx.x == 0 && tid == 0 && !run_init) sh->n_exc[slot_zero] = 0u;
        const unsigned int base = consumed;

        double acc[T::NV];
#pragma unroll
        for (int j = 0; j < T::NV; ++j) acc[j] = 0.0;
        SweepScalars S;
        S.gmax = 0ull; S.eemax = 0ull; S.flags = 0u;
        Held<NF> held;
#pragma unroll
        for (int j = 0; j < NFa; ++j) { held.kv[j] = 0.0; held.pv[j] = 0.0; }
        held.ks = 0.0; held.ps = 0.0;
        const double c2 = 2.0 / (2.0 + P.mot.k), c2c = 2.0 / (2.0 + P.cand.k);
        const double rfac = P.radius / (P.radius + 1.0), inv_radius = 1.0 / P.radius;

        // ---- the sweep: 2 steps per 512-block tile, one residual block per thread and step
        // (the arrays are padded to whole tiles with whatever the allocation held: blocks past the end are replaced
        //  by a harmless constant block as they are read)
        auto load_block = [&](const Stage &st, int j, int index) {
            Loaded X;
            X.p = st.xy[j]; X.u = st.uu[j]; X.a = st.aa[j]; X.d = st.d[j];
            if (index >= D.m) { X.p = make_double2(0.0, 0.0); X.u = X.p; X.a = make_double2(1.0, 0.0); X.d = 1.0; }
            return X;
        };
        Loaded Ln;
        Ln.p = make_double2(0.0, 0.0); Ln.u = Ln.p; Ln.a = make_double2(1.0, 0.0); Ln.d = 1.0;
        if (n_my > 0) {
            const int s = (int)(base % kStages);
            mbar_wait(&full[s], (base / kStages) & 1u);
            Ln = load_block(stages[s], tid, (int)blockIdx.x * kTile + tid);
        }
        for (int k = 0; k < n_my; ++k) {
            const unsigned int g = base + (unsigned)k;
            const int s = (int)(g % kStages);
            const int base_i = ((int)blockIdx.x + k * G) * kTile;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const Loaded L = Ln;
                const int idx = base_i + tid + half * kThreads;
                const bool valid = idx < D.m;
                // fetch the next step's residual block while this one is computed
                if (half == 0) {
                    Ln = load_block(stages[s], tid + kThreads, base_i + tid + kThreads);
                    // this warp has read both halves of the stage: release it (one arrival per warp)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s]);
                } else {
                    // producer: the next tile must be on its way before anybody waits for it (blocks only if a warp
                    // still holds the stage that tile needs); beyond that, refill whatever has been released
                    if (tid == 0) {
                        const unsigned int lim = base + (unsigned)n_my;
                        queue_uses(g + 2u < lim ? g + 2u : lim, base, dx, true);
                        queue_uses(lim, base, dx, false);
                    }
                    if (k + 1 < n_my) {
                        const int s1 = (int)((g + 1u) % kStages);
                        mbar_wait(&full[s1], ((g + 1u) / kStages) & 1u);
                        Ln = load_block(stages[s1], tid, base_i + G * kTile + tid);
                    }
                }
                // rank-1 updates of the previous step (independent of everything below) ...
                accumulate_held<NF>(held, acc);
                // ... and this step's residual / Jacobian chain
                double kv[NFa], sv[NFa], ks, ss;
                if (run_init)
                    init_pixel<NF>(L, valid, idx, P, c2, acc, S, kv, sv, ks, ss, n_exc, &sh->exc_overflow, elist_p, exc_cap);
                else
                    fused_pixel<NF>(L, valid, idx, P, c2, c2c, rfac, inv_radius, acc, S, dcand, kv, sv, ks, ss, n_exc, &sh->exc_overflow,
                                    elist_p, exc_cap);
                if (NF > 0) {
#pragma unroll
                    for (int j = 0; j < NF; ++j) { held.kv[j] = kv[j]; held.pv[j] = __shfl_xor_sync(0xffffffffu, sv[j], 1); }
                    held.ks = ks;
                    held.ps = __shfl_xor_sync(0xffffffffu, ss, 1);
                }
            }
        }
        accumulate_held<NF>(held, acc);
        consumed += (unsigned)n_my;
        acc[T::iGMAX] = __longlong_as_double((long long)S.gmax);
        acc[T::iEEMAX] = __longlong_as_double((long long)S.eemax);
        acc[T::iBADRES] = (S.flags & 1u) ? 1.0 : 0.0;
        acc[T::iBAD] = (S.flags & 2u) ? 1.0 : 0.0;
        acc[T::iBADSTEP] = (S.flags & 4u) ? 1.0 : 0.0;
        // the candidate depths written above are read by TMA in the next phase: order them for the async proxy
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncthreads();
        const unsigned long long t_loop = t_begin ? globaltimer() : 0ull;
        double *row = A_.partials + ((size_t)(gen & 1u) * G + blockIdx.x) * A::NV;
        cta_reduce_roles<NF, LD>(acc, part, row);
        // ---- arrive; then queue the next phase's first tiles (depth: from the buffer an ACCEPTED step makes current)
        if (tid == 0) {
            __threadfence();
            atomicAdd(&sh->arrive, 1u);
            if (t_begin) {
                const unsigned long long t2 = globaltimer();
                atomicAdd(&sh->t_phase[run_init ? 4 : 7], t_loop - t_begin); atomicAdd(&sh->t_phase[run_init ? 5 : 8], t2 - t_loop);
            }
            queue_uses(consumed + (unsigned)pre, consumed, run_init ? dx : dcand, true);
            const unsigned long long t0 = globaltimer();
            const unsigned int target = (gen + 1u) * (unsigned)G;
            int err = 0;
            while ((int)(ld_acquire(&sh->arrive) - target) < 0) {
                __nanosleep(20);
                if (globaltimer() - t0 > kWatchdogNs) { err = 1; sh->error = 1; break; }
            }
            if (!err && __ldcg(&sh->error)) err = 1;
            s_flag[5] = err;
        }
        __syncthreads();
        if (s_flag[5]) {                                          // watchdog: give up, but leave no bulk copy in flight
            for (unsigned int u = consumed; u < consumed + (unsigned)pre; ++u) mbar_wait(&full[(int)(u % kStages)], (u / kStages) & 1u);
            if (tid == 0) P.error = 1;
            __syncthreads();
            break;
        }
        const unsigned long long t_ctl = t_begin ? globaltimer() : 0ull;

        // ---- every CTA: sum the G rows in a fixed order (warp w: rows w, w+8, ...; lanes: columns)
        {
            constexpr int nv = A::NV, ns = A::NS;
            const double *rows = A_.partials + (size_t)(gen & 1u) * G * A::NV;
            if (tid < kExcSlots) s_ne[tid] = __ldcg(&sh->n_exc[tid]);
            constexpr int kRowsPerWarp = (kNumSMsB200 + kWarps - 1) / kWarps;      // 19
            double v[3] = {0.0, 0.0, 0.0};
            if (G <= kNumSMsB200) {
                double t[kRowsPerWarp][3];
#pragma unroll
                for (int u = 0; u < kRowsPerWarp; ++u) {
                    const int b = u * kWarps + warp;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int j = lane + 32 * c;
                        t[u][c] = (b < G && j < nv) ? __ldcg(rows + (size_t)b * A::NV + j) : 0.0;
                    }
                }
#pragma unroll
                for (int u = 0; u < kRowsPerWarp; ++u)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int j = lane + 32 * c;
                        v[c] = (j < ns) ? v[c] + t[u][c] : fmax(v[c], t[u][c]);
                    }
            } else {
                for (int b = warp; b < G; b += kWarps)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int j = lane + 32 * c;
                        const double x = (j < nv) ? __ldcg(rows + (size_t)b * A::NV + j) : 0.0;
                        v[c] = (j < ns) ? v[c] + x : fmax(v[c], x);
                    }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) { const int j = lane + 32 * c; if (j < nv) part[warp][j] = v[c]; }
            __syncthreads();
            if (tid < nv) {
                double x = part[0][tid];
                if (tid < ns) for (int w = 1; w < kWarps; ++w) x += part[w][tid];
                else          for (int w = 1; w < kWarps; ++w) x = fmax(x, part[w][tid]);
                fin[tid] = x;
            }
            __syncthreads();
        }
        const unsigned long long t_fin = t_begin ? globaltimer() : 0ull;
        // ---- FUSED: judge the candidate first
        if (tid == 0) {
            int accepted = run_init ? 1 : 0;
            LmNext nx = LM_RUN_A;
            if (!run_init) {
                CandSums c;
                c.mcc = fin[A::oMCC]; c.step_sq = fin[A::oSTEP]; c.cand_cost = 0.5 * fin[0];
                c.bad_step = fin[A::iBADSTEP]; c.bad_cand = fin[A::iBADRES];
                nx = s_ctl.on_candidate(c);
                accepted = (nx == LM_RUN_A) ? 1 : 0;
            }
            s_flag[1] = (int)nx;
            s_flag[3] = accepted;
            const int cur = run_init ? slot_cur : (accepted ? slot_spec : slot_cur);
            s_flag[4] = cur;
            const unsigned int ne = s_ne[cur];
            s_flag[2] = (int)(ne < exc_cap ? ne : exc_cap);
        }
        __syncthreads();
        const bool accepted = s_flag[3] != 0;
        if (!run_init) {
            // the list that is not current any more is cleared during the next phase and reused after it
            const int dead = accepted ? slot_cur : slot_spec;
            slot_cur = s_flag[4];
            slot_spec = slot_zero;
            slot_zero = dead;
        }
        if (accepted) {
            // the evaluation sums of this pass describe the (new) current point: EvalSums in place
            if (tid < kTri) {
                s_ctl.ev.G1[tid] = (tid < A::TRI) ? fin[A::oG1 + (tid < A::TRI ? tid : 0)] : 0.0;
                s_ctl.ev.G2[tid] = (tid < A::TRI) ? fin[A::oG2 + (tid < A::TRI ? tid : 0)] : 0.0;
            }
            if (tid < kMaxNF) {
                s_ctl.ev.h1[tid] = (tid < NF) ? fin[A::oH1 + (tid < NF ? tid : 0)] : 0.0;
                s_ctl.ev.h2[tid] = (tid < NF) ? fin[A::oH2 + (tid < NF ? tid : 0)] : 0.0;
            }
            if (tid == 32) {
                s_ctl.ev.cost = 0.5 * fin[0]; s_ctl.ev.sumsq_d = fin[1]; s_ctl.ev.gmax_e = fin[A::iGMAX];
                s_ctl.ev.bad = fin[A::iBAD]; s_ctl.ev.ee_max = fin[A::iEEMAX];
            }
            __syncthreads();
        }
        // the pixel ring may hold prefetched tiles: the sort keys of the (rare) listed pixels live in the row scratch
        // of the last stage only when the ring is idle, so listed pixels first drain the ring (see below)
        bool xsorted = false;
        unsigned long long *xkeys = reinterpret_cast<unsigned long long *>(smem_raw);
        bool drained = false;
        if constexpr (NF > 0) if (s_flag[2] > 0) {
            // listed pixels: the controller needs scratch for their sort keys -- give up the prefetched tiles
            for (unsigned int u = consumed; u < consumed + (unsigned)pre; ++u) {
                const int s = (int)(u % kStages);
                mbar_wait(&full[s], (u / kStages) & 1u);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
            }
            consumed += (unsigned)pre;
            drained = true;
            __syncthreads();
            xsorted = sort_exceptions(exc + (size_t)s_flag[4] * exc_cap, s_flag[2], xkeys, 16384, tid);
        }
        if (accepted) {
            if constexpr (NF > 0) if (s_flag[2] > 0) {
                const ExcEntry *cur_exc = exc + (size_t)s_flag[4] * exc_cap;
                double a[kExcVals];
#pragma unroll
                for (int j = 0; j < kExcVals; ++j) a[j] = 0.0;
                for (int k = tid; k < s_flag[2]; k += kThreads) {
                    const int slot = xsorted ? (int)(unsigned int)(xkeys[k] & 0xffffffffull) : k;
                    ExcEntry X;
                    for (int w = 0; w < (int)(sizeof(ExcEntry) / sizeof(double)); ++w)
                        reinterpret_cast<double *>(&X)[w] = __ldcg(reinterpret_cast<const double *>(cur_exc + slot) + w);
                    int t = 0;
#pragma unroll
                    for (int j = 0; j < NF; ++j) {
                        a[kTri + j] += fma(X.F0[j], X.r0, X.F1[j] * X.r1);
#pragma unroll
                        for (int c = j; c < NF; ++c, ++t) a[t] += fma(X.F0[j], X.F0[c], X.F1[j] * X.F1[c]);
                    }
                }
                cta_reduce_sums<kExcVals, LD>(a, part, fin);
                if (tid < A::TRI) s_ctl.ev.G1[tid] += fin[tid];
                if (tid < NF) s_ctl.ev.h1[tid] += fin[kTri + tid];
                __syncthreads();
            }
            if (warp == 0) {
                const int nx = ctl_on_eval<NF>(s_ctl);
                if (lane == 0) s_flag[1] = nx;
            }
            __syncthreads();
        }
        // ---- (re)solve at the current radius; the clamped-pixel correction is summed by the whole CTA
        while (s_flag[1] == (int)LM_SOLVE) {
            const int ne = s_flag[2];
            if constexpr (NF > 0) if (ne > 0) {
                const ExcEntry *cur_exc = exc + (size_t)s_flag[4] * exc_cap;
                const double R = s_ctl.radius, lo = s_ctl.opt.min_lm_diagonal, hi = s_ctl.opt.max_lm_diagonal;
                double a[kExcVals];
#pragma unroll
                for (int j = 0; j < kExcVals; ++j) a[j] = 0.0;
                for (int k = tid; k < ne; k += kThreads) {
                    const int slot = xsorted ? (int)(unsigned int)(xkeys[k] & 0xffffffffull) : k;
                    ExcEntry X;
                    for (int w = 0; w < (int)(sizeof(ExcEntry) / sizeof(double)); ++w)
                        reinterpret_cast<double *>(&X)[w] = __ldcg(reinterpret_cast<const double *>(cur_exc + slot) + w);
                    const double q = X.se2 / (X.ees + fmin(fmax(X.ees, lo), hi) / R);
                    const double er = fma(X.e0, X.r0, X.e1 * X.r1);
                    double fe[NFa];
#pragma unroll
                    for (int j = 0; j < NF; ++j) fe[j] = fma(X.F0[j], X.e0, X.F1[j] * X.e1);
                    int t = 0;
#pragma unroll
                    for (int j = 0; j < NF; ++j) {
                        const double qf = q * fe[j];
                        a[kTri + j] = fma(qf, er, a[kTri + j]);
#pragma unroll
                        for (int c = j; c < NF; ++c, ++t) a[t] = fma(qf, fe[c], a[t]);
                    }
                }
                cta_reduce_sums<kExcVals, LD>(a, part, fin);
                if (tid < kTri) s_exc.S[tid] = fin[tid];
                if (tid < kMaxNF) s_exc.rhs[tid] = fin[kTri + tid];
                __syncthreads();
            }
            if (warp == 0) {
                const int nx = ctl_solve<NF>(s_ctl, (NF > 0 && ne > 0) ? &s_exc : nullptr, s_L, s_y);
                if (lane == 0) s_flag[1] = nx;
            }
            __syncthreads();
        }
        // ---- next phase parameters (every CTA writes its own copy)
        const bool which_changed = !run_init && accepted;
        if (tid == 0) {
            const LmNext nx = (LmNext)s_flag[1];                     // LM_RUN_B (another fused pass) or LM_DONE
            if (which_changed) P.which_x ^= 1;                       // the candidate became x
            Motion mo = P.base, ca = P.base;
            if (NF >= 6) for (int j = 0; j < 3; ++j) {
                mo.v[j] = s_ctl.f[j]; mo.w[j] = s_ctl.f[3 + j];
                ca.v[j] = s_ctl.f[j] + s_ctl.delta_f[j]; ca.w[j] = s_ctl.f[3 + j] + s_ctl.delta_f[3 + j];
            }
            if (NF == 7) { mo.k = s_ctl.f[6]; ca.k = s_ctl.f[6] + s_ctl.delta_f[6]; }
            if (nx == LM_DONE && s_ctl.termination == RSDSFM_FAILURE) mo = P.base;   // Ceres restores the start values
            P.mot = mo; P.cand = ca;
            for (int j = 0; j < kMaxNF; ++j) P.delta_f[j] = s_ctl.delta_f[j];
            P.radius = s_ctl.radius;
            P.ee_fast_min = s_ctl.ee_fast_min;
            P.first = 0;
            P.next = (int)nx;
            if (t_begin) {
                const unsigned long long t_end = globaltimer();
                atomicAdd(&sh->t_phase[run_init ? 10 : 11], t_end - t_fin);
                atomicAdd(&sh->t_phase[run_init ? 6 : 9], t_end - t_ctl);
                atomicAdd(&sh->t_phase[run_init ? 0 : 2], t_end - t_begin); atomicAdd(&sh->t_phase[run_init ? 1 : 3], 1ull);
            }
        }
        __syncthreads();
        gen++;
        // ---- the depth prefetch assumed "accepted" (or INIT): anything else reloads the first tiles
        const bool spec_ok = run_init || accepted;
        if (P.next != LM_DONE && (!spec_ok || drained)) {
            if (!drained) {
                for (unsigned int u = consumed; u < consumed + (unsigned)pre; ++u) {
                    const int s = (int)(u % kStages);
                    mbar_wait(&full[s], (u / kStages) & 1u);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s]);
                }
                consumed += (unsigned)pre;
            }
            if (tid == 0) queue_uses(consumed + (unsigned)pre, consumed, P.which_x ? d1 : d0, true);
        } else if (P.next == LM_DONE && !drained) {
            // leave no bulk copy in flight when the CTA exits
            for (unsigned int u = consumed; u < consumed + (unsigned)pre; ++u) mbar_wait(&full[(int)(u % kStages)], (u / kStages) & 1u);
        }
    }

    // ---- the result: CTA 0 publishes the controller state and the final motion
    if (blockIdx.x == 0) {
        __syncthreads();
        for (int w = tid; w < (int)(sizeof(LmController) / sizeof(int)); w += kThreads)
            reinterpret_cast<int *>(&sh->ctl)[w] = reinterpret_cast<const int *>(&s_ctl)[w];
        if (tid < (int)(sizeof(Bcast) / sizeof(int)))
            reinterpret_cast<int *>(&sh->bc)[tid] = reinterpret_cast<const int *>(&P)[tid];
    }
    // ---- epilogue: write the result (z = 1/d for a9, d for a8).  On FAILURE Ceres restores the
    // start values (solver.cc Minimize): 1/z_in for a9 (double reciprocal, :213/:247), 1.0 for a8.
    const bool failed = (s_ctl.termination == RSDSFM_FAILURE) || P.error;
    const double *dfin = P.which_x ? d1 : d0;
    double zs[1] = {0.0}, zm[2] = {-INFINITY, -INFINITY};
    for (int i = blockIdx.x * kThreads + tid; i < D.m; i += G * kThreads) {
        double dv;
        if (failed) dv = A_.z_in ? 1.0 / A_.z_in[(size_t)i * A_.z_stride] : 1.0;
        else dv = dfin[i];
        const double o = A_.invert_out ? 1.0 / dv : dv;
        A_.out[i] = o;
        zs[0] += o; zm[0] = fmax(zm[0], o); zm[1] = fmax(zm[1], -o);
    }
    if (A_.zstats) block_reduce_store<1, 2>(zs, zm, A_.zstats);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// RSDSFM_LM_VARIANT=1 selects the first-generation kernel (A/B measurements); default: k_lm_solve
static int lm_variant()
{
    static const int v = [] { const char *e = getenv("RSDSFM_LM_VARIANT"); return e ? atoi(e) : 2; }();
    return v;
}

template <int NF>
static int launch_persistent(rsdsfm_ctx *ctx, RefineData D, double *d0, double *d1, LmShared *sh, double *partials,
                             ExcEntry *exc, unsigned int exc_cap, const double *z_in, int z_stride, double *out,
                             int invert_out, double *zstats, int grid)
{
    const size_t smem = sizeof(Stage) * (size_t)kStages;
    if (ctx->profile) cudaEventRecord(ctx->pe0[ctx->io_slot], ctx->stream);
    if (lm_variant() == 1) {
        RS_CUDA(ctx, cudaFuncSetAttribute(k_lm_persistent<NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        void *args[] = {&D, &d0, &d1, &sh, &partials, &exc, &exc_cap, &z_in, &z_stride, &out, &invert_out, &zstats};
        RS_CUDA(ctx, cudaLaunchCooperativeKernel((void *)k_lm_persistent<NF>, dim3(grid), dim3(kThreads), args, smem, ctx->stream));
    } else {
        RS_CUDA(ctx, cudaFuncSetAttribute(k_lm_solve<NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SolveArgs a{D, d0, d1, sh, partials, exc, exc_cap, z_in, z_stride, out, invert_out, zstats};
        void *args[] = {&a};
        RS_CUDA(ctx, cudaLaunchCooperativeKernel((void *)k_lm_solve<NF>, dim3(grid), dim3(kThreads), args, smem, ctx->stream));
    }
    if (ctx->profile) cudaEventRecord(ctx->pe1[ctx->io_slot], ctx->stream);
    ctx->launches++;
    return RSDSFM_OK;
}

// The initial control block travels as a kernel parameter and the final one is written straight
// into pinned host memory by a kernel: neither transfer queues behind the large uploads /
// downloads that a pipelined sequence keeps on the copy engines.
static_assert(sizeof(LmShared) <= 3840 && sizeof(LmShared) % 4 == 0, "control block must fit the kernel parameter space");

__global__ void __launch_bounds__(256) k_lm_begin(LmShared *sh, const __grid_constant__ LmShared init, int keep_input_flag)
{
    // nonfinite_input is the LAST field: with keep_input_flag the gather kernel's verdict survives
    const int words = (int)((keep_input_flag ? offsetof(LmShared, nonfinite_input) : sizeof(LmShared)) / 4);
    const unsigned int *src = reinterpret_cast<const unsigned int *>(&init);
    unsigned int *dst = reinterpret_cast<unsigned int *>(sh);
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    // "gather found a non-finite start depth" => the FAILURE Ceres reports before evaluating anything
    if (threadIdx.x == 0 && keep_input_flag && sh->nonfinite_input) {
        sh->bc.next = LM_DONE;
        sh->ctl.termination = RSDSFM_FAILURE;
        sh->ctl.reason = RSDSFM_REASON_NONFINITE_INPUT;
    }
}

__global__ void __launch_bounds__(256) k_lm_readback(const LmShared *sh, const double *stats8, LmShared *host_block,
                                                     double *host_stats8)
{
    const unsigned int *src = reinterpret_cast<const unsigned int *>(sh);
    unsigned int *dst = reinterpret_cast<unsigned int *>(host_block);            // mapped pinned memory (UVA)
    for (int i = threadIdx.x; i < (int)(sizeof(LmShared) / 4); i += blockDim.x) dst[i] = src[i];
    if (stats8 && threadIdx.x < 8) host_stats8[threadIdx.x] = stats8[threadIdx.x];
}

int lm_grid_size(const rsdsfm_ctx *ctx) { return (ctx->lm_grid > 0 && ctx->lm_grid < ctx->num_sms) ? ctx->lm_grid : ctx->num_sms; }

// Initial capacity (entries) of each clamped-pixel list; an overflow is detected by the kernel and the
// solve is repeated with room for every residual block.  RSDSFM_EXC_CAP overrides the default so that
// the tests can drive the overflow path with a handful of clamped pixels.
static int min_exc_cap()
{
    const char *e = getenv("RSDSFM_EXC_CAP");
    const int v = e ? atoi(e) : 0;
    return v > 0 ? v : 4096;
}

// Queues one LM solve on the context's stream (no host synchronisation).  The control block
// (ctx->lm_shared) receives the result; lm_collect() reads it back.
static int lm_solve_async(rsdsfm_ctx *ctx, const RefineData &D, double *d0, double *d1, int nf, const Motion &mot0,
                          const rsdsfm_lm_options &opt, const double *z_in, int z_stride, double *out, int invert_out,
                          bool keep_input_flag, double *zstats = nullptr)
{
    RS_TRY(ensure(ctx, ctx->lm_shared, sizeof(LmShared)));
    LmShared *sh = (LmShared *)ctx->lm_shared.p;
    const int grid = lm_grid_size(ctx);                 // one persistent CTA per SM (half of them in a two-lane sequence)
    const int nv = (nf == 0) ? Acc<0>::NV : (nf == 6 ? Acc<6>::NV : Acc<7>::NV);
    RS_TRY(ensure(ctx, ctx->partials, sizeof(double) * 2 * (size_t)grid * nv));   // rows of even / odd phases
    if (ctx->exc_cap < min_exc_cap()) ctx->exc_cap = min_exc_cap();
    RS_TRY(ensure(ctx, ctx->exc, sizeof(ExcEntry) * kExcSlots * (size_t)ctx->exc_cap));   // current + speculative + being cleared
    static_assert(sizeof(LmShared) <= 8192 - 256, "pinned slot layout (common.cuh)");

    // initial control block: built on the host, passed by value to k_lm_begin
    LmShared init_block;
    LmShared *h = &init_block;
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
    k_lm_begin<<<1, 256, 0, ctx->stream>>>(sh, *h, keep_input_flag ? 1 : 0);
    ctx->launches++;
    double *partials = (double *)ctx->partials.p;
    ExcEntry *exc = (ExcEntry *)ctx->exc.p;
    const unsigned int cap = (unsigned int)ctx->exc_cap;
    if (nf == 0) return launch_persistent<0>(ctx, D, d0, d1, sh, partials, exc, cap, z_in, z_stride, out, invert_out, zstats, grid);
    if (nf == 6) return launch_persistent<6>(ctx, D, d0, d1, sh, partials, exc, cap, z_in, z_stride, out, invert_out, zstats, grid);
    return launch_persistent<7>(ctx, D, d0, d1, sh, partials, exc, cap, z_in, z_stride, out, invert_out, zstats, grid);
}

// Queues the read-back of the control block (and, if given, of the 8 depth statistics) into the
// current I/O slot's pinned area.
int lm_collect_enqueue(rsdsfm_ctx *ctx, const double *stats_dev)
{
    k_lm_readback<<<1, 256, 0, ctx->stream>>>((const LmShared *)ctx->lm_shared.p, stats_dev, (LmShared *)pinned_lm_result(ctx),
                                              pinned_stats(ctx));
    ctx->launches++;
    RS_CUDA(ctx, cudaGetLastError());
    return RSDSFM_OK;
}

// Parses the control block of the current I/O slot once its read-back has completed.
// RSDSFM_ERR_INTERNAL when the in-kernel watchdog tripped; *overflow: the exception list was too
// small (ctx->exc_cap was raised: run the solve again).
int lm_collect_finish(rsdsfm_ctx *ctx, int nf, int m, Motion *mot, rsdsfm_lm_summary *summary, bool *overflow)
{
    LmShared *h = (LmShared *)pinned_lm_result(ctx);
    if (h->error) return fail(ctx, RSDSFM_ERR_INTERNAL, "LM kernel: grid barrier watchdog tripped");
    if (h->nonfinite_input & 2)
        return fail(ctx, RSDSFM_ERR_ARG, "compact input: n / m do not match the flow field and the consensus mask");
    *overflow = h->exc_overflow != 0;
    if (*overflow) { if (ctx->exc_cap < m + 1024) ctx->exc_cap = m + 1024; return RSDSFM_OK; }
    float kms = 0.f;
    if (ctx->profile) cudaEventElapsedTime(&kms, ctx->pe0[ctx->io_slot], ctx->pe1[ctx->io_slot]);
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
        for (int j = 0; j < 8; ++j) ctx->prof_detail[j] += (double)h->t_phase[4 + j] * 1e-6;
    }
    if (mot && h->ctl.termination != RSDSFM_FAILURE) {
        if (nf >= 6) for (int j = 0; j < 3; ++j) { mot->v[j] = h->ctl.f[j]; mot->w[j] = h->ctl.f[3 + j]; }
        if (nf == 7) mot->k = h->ctl.f[6];
    }
    return RSDSFM_OK;
}

// Reads the control block back (synchronises the stream).
int lm_collect(rsdsfm_ctx *ctx, int nf, int m, Motion *mot, rsdsfm_lm_summary *summary, bool *overflow)
{
    RS_TRY(lm_collect_enqueue(ctx, nullptr));
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return lm_collect_finish(ctx, nf, m, mot, summary, overflow);
}

// Device address of the refined motion (v[3], w[3], k) inside the control block, for the
// rectification stage that follows on the same stream.
const double *lm_motion_device(rsdsfm_ctx *ctx)
{
    return reinterpret_cast<const double *>((const char *)ctx->lm_shared.p + offsetof(LmShared, bc) + offsetof(Bcast, mot));
}

static int ensure_lm_buffers(rsdsfm_ctx *ctx, size_t mm)
{
    const size_t tiles = (mm + kTile - 1) / kTile;             // everything is padded to whole tiles
    RS_TRY(ensure(ctx, ctx->pix, sizeof(double2) * 3 * kTile * tiles));
    RS_TRY(ensure(ctx, ctx->dA, sizeof(double) * kTile * tiles));
    RS_TRY(ensure(ctx, ctx->dB, sizeof(double) * kTile * tiles));
    RS_TRY(ensure(ctx, ctx->lm_shared, sizeof(LmShared)));
    return RSDSFM_OK;
}

// Pre-sizes every solver buffer for up to m residual blocks so that a pipelined sequence does not
// reallocate (and therefore synchronise) between pairs.
int lm_reserve(rsdsfm_ctx *ctx, int m)
{
    RS_TRY(ensure_lm_buffers(ctx, (size_t)(m > 0 ? m : 1)));
    RS_TRY(ensure(ctx, ctx->partials, sizeof(double) * 2 * (size_t)ctx->num_sms * Acc<7>::NV));
    if (ctx->exc_cap < min_exc_cap()) ctx->exc_cap = min_exc_cap();
    return ensure(ctx, ctx->exc, sizeof(ExcEntry) * kExcSlots * (size_t)ctx->exc_cap);
}

// a9 on device pointers: queues gather + solve on the stream, no synchronisation.
int refine_async(rsdsfm_ctx *ctx, const double *flow, const double *inliers3, const double *alpha,
                 const double *alpha_k, int m, const double *v, const double *w, double k, int const_acc,
                 const int32_t *flow_index, const rsdsfm_lm_options *opts, double *z_out, double *zstats)
{
    rsdsfm_lm_options o;
    if (opts) o = *opts; else rsdsfm_lm_default_options(&o);
    const size_t mm = (size_t)m;
    RS_TRY(ensure_lm_buffers(ctx, mm));
    double2 *blk = (double2 *)ctx->pix.p;
    double *d0 = (double *)ctx->dA.p, *d1 = (double *)ctx->dB.p;
    LmShared *sh = (LmShared *)ctx->lm_shared.p;
    RS_CUDA(ctx, cudaMemsetAsync(&sh->nonfinite_input, 0, sizeof(int), ctx->stream));
    k_refine_gather<<<grid_for(ctx, m, 8), kThreads, 0, ctx->stream>>>(flow, inliers3, alpha, alpha_k, flow_index, m, blk, d0, sh);
    ctx->launches++;
    RefineData D{blk, m};
    Motion mot;
    for (int j = 0; j < 3; ++j) { mot.v[j] = v[j]; mot.w[j] = w[j]; }
    mot.k = k;
    return lm_solve_async(ctx, D, d0, d1, const_acc ? 7 : 6, mot, o, inliers3 + 2, 3, z_out, 1, true, zstats);
}

// The solver's input buffers for m residual blocks, for a producer that fills them itself
// (preproc.cu: compact_build_device): tile-blocked blk, the start inverse depths d0, and the
// device flag the producer raises for bad input (1: non-finite start depth, 2: count mismatch).
int lm_input_buffers(rsdsfm_ctx *ctx, int m, void **blk, double **d0, int **input_flag)
{
    RS_TRY(ensure_lm_buffers(ctx, (size_t)(m > 0 ? m : 1)));
    LmShared *sh = (LmShared *)ctx->lm_shared.p;
    RS_CUDA(ctx, cudaMemsetAsync(&sh->nonfinite_input, 0, sizeof(int), ctx->stream));
    *blk = ctx->pix.p; *d0 = (double *)ctx->dA.p; *input_flag = &sh->nonfinite_input;
    return RSDSFM_OK;
}

// a9 on buffers filled through lm_input_buffers: queues the solve, no synchronisation.
// z_in[m]: the start depths (restored on FAILURE, like Ceres restores its parameter blocks).
int refine_prepared_async(rsdsfm_ctx *ctx, int m, const double *v, const double *w, double k, int const_acc,
                          const rsdsfm_lm_options *opts, const double *z_in, double *z_out, double *zstats)
{
    rsdsfm_lm_options o;
    if (opts) o = *opts; else rsdsfm_lm_default_options(&o);
    RefineData D{(const double2 *)ctx->pix.p, m};
    Motion mot;
    for (int j = 0; j < 3; ++j) { mot.v[j] = v[j]; mot.w[j] = w[j]; }
    mot.k = k;
    return lm_solve_async(ctx, D, (double *)ctx->dA.p, (double *)ctx->dB.p, const_acc ? 7 : 6, mot, o, z_in, 1, z_out, 1, true, zstats);
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
        RS_TRY(refine_async(ctx, flow, inliers3, alpha, alpha_k, m, v, w, *k, const_acc, flow_index, opts, z_out, nullptr));
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

// a8 on device pointers.
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
    RS_TRY(ensure_lm_buffers(ctx, nn));
    double2 *blk = (double2 *)ctx->pix.p;
    double *d0 = (double *)ctx->dA.p, *d1 = (double *)ctx->dB.p;
    k_depth_gather<<<grid_for(ctx, n, 8), kThreads, 0, ctx->stream>>>(coord, flow, alpha, alpha_k, n, blk, d0);
    ctx->launches++;
    RefineData D{blk, n};
    Motion mot;
    for (int j = 0; j < 3; ++j) { mot.v[j] = v[j]; mot.w[j] = w[j]; }
    mot.k = k;
    RS_TRY(lm_solve_async(ctx, D, d0, d1, 0, mot, o, nullptr, 1, inv_depth, 0, false));
    bool overflow = false;
    return lm_collect(ctx, 0, n, nullptr, summary, &overflow);
}

}  // namespace rsdsfm
