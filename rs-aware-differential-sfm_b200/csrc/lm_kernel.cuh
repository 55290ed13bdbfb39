// lm_kernel.cuh -- k_lm_solve<NF>: the persistent cooperative Levenberg-Marquardt kernel behind
//   a8  nonlinear_refinement::estimateInverseDepths  (nonlinearRefinement.cc:109-180)  NF = 0
//   a9  nonlinear_refinement::nonLinearRefinement    (nonlinearRefinement.cc:183-252)  NF = 6 | 7
// (included by refine.cu only; the host side lives there).
//
// One CTA per SM stays resident for the whole solve and loops over LM phases:
//   INIT phase (once): residual + analytic Jacobian at the start point, closed-form 1x1 Schur
//       elimination of the pixel's inverse depth, FP64 accumulation of the radius-independent
//       factors G1, G2, h1, h2 (two rank-1 updates per pixel), cost, |x|^2, max gradient.
//   FUSED phase (one per LM iteration): depth back-substitution of the candidate step at x
//       (candidate depth, model cost change, |step|^2) and, in the same sweep, the evaluation of
//       the next iteration's sums AT THE CANDIDATE, speculatively.  A rejected step re-solves from
//       the stored factors at the smaller radius: no sweep at all.
//
// What shapes the code (sm_100a, measured: FP64 issue is the binding resource -- one warp-DFMA per 2
// cycles per SM sub-partition, 8 cycles dependent latency, two warps per sub-partition at ~250 registers):
//  * STRAIGHT-LINE SWEEP BODY.  One residual block per thread and step, no branch between the
//    shared-memory loads and the rank-1 updates: padded blocks are all-zero records that contribute
//    nothing by construction, range tests are integer compares on the bit patterns, and everything
//    rare (LM diagonal clamped at the focus of expansion, non-finite values) is decided by ONE warp
//    vote per step, which sends the warp through an out-of-line exact path.
//  * TMA RING.  256-block tiles (12 KB + 2 KB bulk copies, cp.async.bulk + mbarrier complete_tx), 14
//    stages per SM, full/empty mbarriers; consumers release a stage with one arrival per warp, and the
//    warps take turns at refilling released stages (see sweep()).
//  * REPLICATED CONTROLLER.  Every CTA publishes one row of partial sums, arrives once on a grid
//    counter, reads all rows, sums them in the same fixed order and runs the same Ceres logic on its
//    own copy of the state: bit-identical decisions, no serial publish/release step.  The next phase's
//    first tiles are prefetched across the barrier (depths from the buffer an accepted step makes current).
//  * CTA REDUCTION BY TRANSPOSITION.  The 39 per-thread sums are reduced with a halving butterfly
//    (each shuffle level sends half of the remaining values): 38 instead of 156 64-bit shuffles.
#pragma once

#include "common.cuh"
#include "lm_controller.h"
#include "lm_layout.h"

namespace rsdsfm {

struct RefineData {
    const double2 *blk;      // [num_tiles][3][kTile]
    int m;
};

struct ExcEntry {            // a pixel whose LM diagonal is (possibly) clamped, or whose e-column is degenerate:
    double ees, se2;         // s_e^2 e^Te, s_e^2     (the whole pixel is handled by the controller)
    double r0, r1, e0, e1;   // residual and depth column
    double F0[kMaxNF], F1[kMaxNF];   // the two Jacobian rows of the free motion parameters
    double key;              // residual-block index: the controller sums the list in ascending key order, so
                             // the result does not depend on the order the atomics handed out the slots
};

// First phase parameters: written by k_lm_begin, read by every CTA at kernel start; rewritten by CTA 0 at the end.
struct Bcast {
    int next, which_x, first, cur_list;
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
    // device timing (globaltimer ns): [0] INIT total, [1] phases, [2] FUSED total, [3] phases,
    // [4..6] INIT pixel loop / CTA reduce / controller, [7..9] same for FUSED, [10], [11] controller logic only
    unsigned long long t_phase[12];
    unsigned long long t_abs[4];   // globaltimer of CTA 0: kernel's first / last instruction, first phase's start, last phase's end (RSDSFM_TRACE)
    // ---- grid synchronisation
    unsigned int arrive, generation;
    unsigned int n_exc[4], exc_overflow, pad1;   // exception lists: current / speculative / being cleared
    int error;
    int nonfinite_input;     // LAST field: raised by the gather kernel, preserved by the control-block upload
};

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
// The sweep's versions: ONE Newton step (seed 2^-23 -> 2^-45 relative).  The reciprocal only scales a depth STEP and
// the reciprocal square root only normalises a direction whose length cancels up to that factor in the Schur sums:
// an error of 3e-14 moves neither the fixed point nor, within the parity tolerances (1e-6 / 1e-4), the path.
__device__ __forceinline__ double sweep_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return fma(r, fma(-x, r, 1.0), r);
}
__device__ __forceinline__ double sweep_rsqrt(double x)
{
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r * fma(-0.5 * x * r, r, 1.5);
}

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
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
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- the same, on 32-bit shared-window addresses that the sweep computes once (a generic pointer to shared
// memory makes the compiler rebuild the window base -- S2R SR_CgaCtaId, LEA, ... -- at every use)
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ bool mbar_test_s(uint32_t bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_s(uint32_t bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_s(uint32_t dst, const void *src, unsigned bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

// non-negative doubles order like their bit patterns; a NaN or a negative value fails both tests
__device__ __forceinline__ bool bits_in_range(double v, unsigned long long lo, unsigned long long hi)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return b >= lo && b <= hi;
}
__device__ __forceinline__ bool not_finite(double x) { return (__double2hiint(x) & 0x7ff00000) == 0x7ff00000; }
__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
__device__ __forceinline__ unsigned long long dbits(double v) { return (unsigned long long)__double_as_longlong(v); }

// ------------------------------------------------------------------------------------------
// Row split of ONE solve over several GPUs (BASELINE config 4).  Every GPU sweeps its own share of the
// residual blocks with this same kernel; what crosses the GPUs is one row of sums per phase (and the sums
// of the listed pixels, when there are any).  The exchange goes through peer memory (NVLink), without NCCL
// and without leaving the persistent kernel: every value travels with the tag of its exchange in ONE 16-byte
// store -- a reader that sees the tag it waits for also sees the value -- so there is neither a fence nor a
// flag.  CTA 0 of every GPU writes its GPU's totals into the mailbox of every GPU; every CTA then waits
// for the rows of all GPUs in its local mailbox and combines them in GPU order: all CTAs of all GPUs hold
// bit-identical sums and their replicated controllers take identical decisions.
// ------------------------------------------------------------------------------------------
struct __align__(16) Tagged {
    double v;
    unsigned long long tag;
};
constexpr int kMaxPeers = 8;
constexpr int kMailSlots = 16;          // exchanges in flight before a slot is reused (at most ~8 per phase)
constexpr int kMailLd = 96;             // values per row
struct PeerInfo {
    Tagged *mail[kMaxPeers];            // mail[g]: mailbox of GPU g, [kMailSlots][kMaxPeers][kMailLd]; mail[me] is local memory
    int n, me;
    unsigned int epoch;                 // solve number (the same on every GPU): tags are (epoch << 32) | exchange number
};
__device__ __forceinline__ void st_tagged(Tagged *p, double v, unsigned long long tag)
{
    asm volatile("st.relaxed.sys.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(v)), "l"(tag) : "memory");
}
__device__ __forceinline__ Tagged ld_tagged(const Tagged *p)
{
    long long v; unsigned long long t;
    asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(v), "=l"(t) : "l"(p) : "memory");
    Tagged r; r.v = __longlong_as_double(v); r.tag = t;
    return r;
}

struct PhaseParams {           // shared-memory copy of the phase parameters (+ options, start point)
    int next, which_x, first, cur_list;
    Motion mot, cand;
    double delta_f[kMaxNF];
    double radius;
    double ee_fast_min;
    double min_diag, max_diag;
    double ee_first_min, ee_first_max;   // INIT: e^Te inside [min, max]  =>  the LM diagonal of that depth is certainly not clamped
    Motion base;
    int error;
};
static_assert(offsetof(PhaseParams, ee_fast_min) == offsetof(Bcast, ee_fast_min), "PhaseParams must start with Bcast");

constexpr int kStages = 14;                     // tiles in flight per CTA (14 x 14 KB of the 227 KB shared memory)
struct Stage {
    double2 xy[kTile], uu[kTile], aa[kTile];
    double d[kTile];
};
static_assert(sizeof(Stage) == 56 * kTile && kTile == kThreads, "stage layout: one residual block per thread and tile");

struct Loaded {
    double2 p, u, a;
    double d;
};

// ------------------------------------------------------------------------------------------
// Per-thread sums of a sweep: [0] sum r^2 and [1] sum d^2 at the evaluation point, [2] the candidate
// step's model-cost term and [3] |step|^2, then K (packed upper triangle) and H.  The two Schur
// factors are split across lane pairs: even lanes keep K = G1 / H = h1 (the n-direction), odd lanes
// K = G2 / H = h2 (the e-direction); partners swap the half they do not keep with one shuffle per
// value.  A CTA row holds the even-lane sums, the odd-lane sums, and three integer words.
// ------------------------------------------------------------------------------------------
template <int NF>
struct TAcc {
    static constexpr int TRI = NF * (NF + 1) / 2;
    static constexpr int oK = 4, oH = oK + TRI;
    static constexpr int NS = oH + NF;
};
template <int NF>
struct Row {
    static constexpr int NS = TAcc<NF>::NS;
    static constexpr int oGMAX = 2 * NS, oEEMAX = oGMAX + 1, oFLAGS = oGMAX + 2;   // bit patterns / flag word, combined with integer max / or
    static constexpr int NV = 2 * NS + 3;
};
constexpr int kExcVals = kTri + kMaxNF;          // exception sums: S triangle + rhs
constexpr int kRowLd = 96;                       // >= Row<7>::NV (81): three 32-lane column chunks
static_assert(Row<7>::NV + 3 <= kRowLd && kExcVals <= kRowLd, "row scratch");

struct SweepScalars {                           // per-thread non-FP64 accumulators of a sweep
    unsigned long long gmax, eemax;             // bit patterns of max |e^T r|, max e^Te
    unsigned flags;                             // 1: residual not finite, 2: residual or Jacobian not finite, 4: depth step not finite
};

// What one residual block hands to the accumulation: the two projected, 1/|e|-scaled Jacobian vectors of
// the lane-pair scheme (kv / ks: the direction this lane keeps, sv / ss: the direction its partner keeps)
// and the scalar terms.
template <int NF>
struct PixOut {
    double kv[NF > 0 ? NF : 1], sv[NF > 0 ? NF : 1], ks, ss;
    double rr, dd2, mcc, stp;
    unsigned long long gb, eb;
    unsigned flags;
};

// ------------------------------------------------------------------------------------------
// Rare path of the evaluation: a pixel whose LM diagonal may be clamped (|e| ~ 0, focus of
// expansion) or whose values are not finite.  Nothing of its Jacobian is accumulated by the
// thread: the pixel is listed and the controller adds F^TF, F^Tr (radius independent) and
// subtracts q (F^Te)(e^TF), q (F^Te)(e^Tr) (radius dependent) itself.
// ------------------------------------------------------------------------------------------
template <int NF>
__device__ __forceinline__ void ft_rows(double beta, double dbeta, double d, double x, double y, double p0, double p1,
                                        double (&F0)[NF > 0 ? NF : 1], double (&F1)[NF > 0 ? NF : 1])
{   // F = -beta [d A | B | (dbeta/beta) p]  (lm_layout.h)
    if (NF >= 6) {
        F0[0] = -beta * d;        F1[0] = 0.0;
        F0[1] = 0.0;              F1[1] = -beta * d;
        F0[2] = beta * d * x;     F1[2] = beta * d * y;
        F0[3] = beta * (x * y);   F1[3] = beta * fma(y, y, 1.0);
        F0[4] = -beta * fma(x, x, 1.0); F1[4] = -beta * (x * y);
        F0[5] = beta * y;         F1[5] = -beta * x;
    }
    if (NF == 7) { F0[6] = -dbeta * p0; F1[6] = -dbeta * p1; }
}

// Jacobi scale of this pixel's depth column: 1/(1+|e(x0)|), e(x0) evaluated at the start motion.
__device__ __forceinline__ double depth_scale_at_start(const Loaded &L, const Motion &b)
{
    const double beta0 = (2.0 / (2.0 + b.k)) * fma(b.k, L.a.y, L.a.x);
    const double s0 = beta0 * fma(-L.p.x, b.v[2], b.v[0]), s1 = beta0 * fma(-L.p.y, b.v[2], b.v[1]);
    return 1.0 / (1.0 + sqrt(fma(s0, s0, s1 * s1)));
}

// The exact, general treatment of ONE residual block (everything the straight-line body assumes away):
// candidate step with the clamped q, evaluation with the exact clamp test, listing.  Out of line and
// cold: the sweep calls it only when a warp vote found a block that needs it.
template <int NF, bool INIT>
__device__ __noinline__ void pixel_exact(const Loaded *Lp, int index, const PhaseParams *Pp, double *d_cand, unsigned int *n_exc,
                                         unsigned int *overflow, ExcEntry *exc, unsigned int exc_cap, PixOut<NF> *out)
{
    constexpr int NFa = NF > 0 ? NF : 1;
    const Loaded L = *Lp;
    const PhaseParams &P = *Pp;
    PixOut<NF> O;
#pragma unroll
    for (int j = 0; j < NFa; ++j) { O.kv[j] = 0.0; O.sv[j] = 0.0; }
    O.ks = 0.0; O.ss = 0.0; O.mcc = 0.0; O.stp = 0.0; O.flags = 0u;
    const double x = L.p.x, y = L.p.y;
    const double xy = x * y, xx1 = fma(x, x, 1.0), yy1 = fma(y, y, 1.0);
    Motion me = P.mot;                 // evaluation point
    double de = L.d;
    if (!INIT) {
        const double c2 = 2.0 / (2.0 + P.mot.k);
        const double ak = fma(P.mot.k, L.a.y, L.a.x), beta = c2 * ak;
        const double a0 = fma(-x, P.mot.v[2], P.mot.v[0]), a1 = fma(-y, P.mot.v[2], P.mot.v[1]);
        const double b0 = fma(-xy, P.mot.w[0], fma(xx1, P.mot.w[1], -y * P.mot.w[2]));
        const double b1 = fma(-yy1, P.mot.w[0], fma(xy, P.mot.w[1], x * P.mot.w[2]));
        const double p0 = fma(L.d, a0, b0), p1 = fma(L.d, a1, b1);
        const double r0 = fma(-beta, p0, L.u.x), r1 = fma(-beta, p1, L.u.y);
        const double e0 = -beta * a0, e1 = -beta * a1;
        const double ee = fma(e0, e0, e1 * e1);
        // q = s_e^2 / (s_e^2 e^Te + clamp(s_e^2 e^Te)/radius)  ( = radius/((radius+1) e^Te) when not clamped )
        double q;
        if (ee >= P.ee_fast_min && ee <= P.max_diag) {
            q = fast_rcp(ee) * (P.radius / (P.radius + 1.0));
        } else {
            const double se = depth_scale_at_start(L, P.base);
            const double se2 = se * se, ees = ee * se2;
            q = se2 / (ees + fmin(fmax(ees, P.min_diag), P.max_diag) * (1.0 / P.radius));
        }
        double m0 = 0.0, m1 = 0.0;
        if (NF >= 6) {
            const double *df = P.delta_f;
            const double da0 = fma(-x, df[2], df[0]), da1 = fma(-y, df[2], df[1]);
            const double db0 = fma(-xy, df[3], fma(xx1, df[4], -y * df[5]));
            const double db1 = fma(-yy1, df[3], fma(xy, df[4], x * df[5]));
            m0 = -beta * fma(L.d, da0, db0);
            m1 = -beta * fma(L.d, da1, db1);
            if (NF == 7) {
                const double dbk = c2 * fma(-ak, 0.5 * c2, L.a.y) * df[6];
                m0 = fma(-dbk, p0, m0);
                m1 = fma(-dbk, p1, m1);
            }
        }
        const double delta_e = -q * fma(e0, r0 + m0, e1 * (r1 + m1));
        const double j0 = fma(e0, delta_e, m0), j1 = fma(e1, delta_e, m1);
        const double dc = L.d + delta_e;
        const double dd = L.d - dc;
        d_cand[index] = dc;
        O.mcc = fma(j0, fma(0.5, j0, r0), j1 * fma(0.5, j1, r1));
        O.stp = dd * dd;
        if (not_finite(delta_e)) O.flags |= 4u;
        me = P.cand;
        de = dc;
    }
    // ---- evaluation at (me, de)
    const double c2 = 2.0 / (2.0 + me.k);
    const double ak = fma(me.k, L.a.y, L.a.x), beta = c2 * ak;
    const double a0 = fma(-x, me.v[2], me.v[0]), a1 = fma(-y, me.v[2], me.v[1]);
    const double b0 = fma(-xy, me.w[0], fma(xx1, me.w[1], -y * me.w[2]));
    const double b1 = fma(-yy1, me.w[0], fma(xy, me.w[1], x * me.w[2]));
    const double p0 = fma(de, a0, b0), p1 = fma(de, a1, b1);
    const double r0 = fma(-beta, p0, L.u.x), r1 = fma(-beta, p1, L.u.y);
    const double e0 = -beta * a0, e1 = -beta * a1;
    const double ee = fma(e0, e0, e1 * e1);
    const double dbeta = (NF == 7) ? c2 * fma(-ak, 0.5 * c2, L.a.y) : 0.0;
    O.rr = fma(r0, r0, r1 * r1);
    O.dd2 = de * de;
    O.gb = dbits(fabs(fma(e0, r0, e1 * r1)));
    O.eb = dbits(ee);
    if (not_finite(r0 + r1)) O.flags |= 3u;
    if (not_finite(ee)) O.flags |= 2u;
    const bool first = INIT && P.first != 0;
    double se = 1.0;
    bool fast;
    if (first) {
        se = 1.0 / (1.0 + sqrt(ee));
        const double ees = ee * se * se;
        fast = (ees >= P.min_diag && ees <= P.max_diag);
    } else {
        fast = (ee >= P.ee_fast_min && ee <= P.max_diag);
    }
    if (NF > 0) {
        if (fast) {
            const double mu = fast_rsqrt(ee);
            const double c = mu * e0, s = mu * e1;
            const bool e_role = (threadIdx.x & 1) != 0;
            const double mc = e_role ? c : -s, ms = e_role ? s : c;
            double F0[NFa], F1[NFa];
            ft_rows<NF>(beta, dbeta, de, x, y, p0, p1, F0, F1);
#pragma unroll
            for (int j = 0; j < NF; ++j) { O.kv[j] = fma(F0[j], mc, F1[j] * ms); O.sv[j] = fma(F1[j], mc, -(F0[j] * ms)); }
            O.ks = fma(mc, r0, ms * r1);
            O.ss = fma(mc, r1, -(ms * r0));
        } else {
            if (!first) se = depth_scale_at_start(L, P.base);
            double F0[NFa], F1[NFa];
            ft_rows<NF>(beta, dbeta, de, x, y, p0, p1, F0, F1);
            const unsigned int slot = atomicAdd(n_exc, 1u);
            if (slot < exc_cap) {
                ExcEntry X;
                X.ees = ee * se * se; X.se2 = se * se;
                X.r0 = r0; X.r1 = r1; X.e0 = e0; X.e1 = e1; X.key = (double)index;
#pragma unroll
                for (int j = 0; j < kMaxNF; ++j) { X.F0[j] = (j < NF) ? F0[j < NF ? j : 0] : 0.0; X.F1[j] = (j < NF) ? F1[j < NF ? j : 0] : 0.0; }
                exc[slot] = X;
            } else {
                *overflow = 1u;
            }
        }
    }
    *out = O;
}

// ------------------------------------------------------------------------------------------
// The straight-line body.  Uniform quantities of a phase: built once per phase by one thread in shared
// memory; every thread copies them into registers before its sweep.
// ------------------------------------------------------------------------------------------
struct __align__(16) SweepU {
    double v0, v1, v2, w0, w1, w2, k, c2;          // current point x (FUSED) / evaluation point (INIT)
    double kc, c2c;                                  // evaluation point: k, 2/(2+k)
    double dl[6], K1, K2;                            // motion step delta_f; K1 = c2 dk, K2 = c2^2 dk / 2
    double K4, rfac;                                 // c2c^2 / 2;  radius / (radius + 1)
    unsigned long long lo_x, hi_x, lo_c, hi_c;       // "certainly not clamped" ranges of e^Te (bit patterns) at x / at the evaluation point
};

template <int NF>
__device__ __forceinline__ void sweep_uniforms(const PhaseParams *Pp, SweepU *Up)
{
    const PhaseParams &P = *Pp;
    SweepU U;
    const bool init = (P.next == LM_RUN_A);
    const unsigned long long tiny = 0x0010000000000000ull;     // DBL_MIN: e^Te = 0 (padding, degenerate pixels) is never "fast"
    U.v0 = P.mot.v[0]; U.v1 = P.mot.v[1]; U.v2 = P.mot.v[2];
    U.w0 = P.mot.w[0]; U.w1 = P.mot.w[1]; U.w2 = P.mot.w[2];
    U.k = P.mot.k; U.c2 = 2.0 / (2.0 + P.mot.k);
    if (init) {
        U.kc = U.k; U.c2c = U.c2;
        U.lo_c = umax64(dbits(P.first ? P.ee_first_min : P.ee_fast_min), tiny);
        U.hi_c = dbits(P.first ? P.ee_first_max : P.max_diag);
        U.lo_x = U.lo_c; U.hi_x = U.hi_c;
        U.rfac = 0.0; U.K1 = 0.0; U.K2 = 0.0;
        for (int j = 0; j < 6; ++j) U.dl[j] = 0.0;
    } else {
        U.kc = P.cand.k; U.c2c = 2.0 / (2.0 + P.cand.k);
        U.lo_x = umax64(dbits(P.ee_fast_min), tiny); U.hi_x = dbits(P.max_diag);
        U.lo_c = U.lo_x; U.hi_c = U.hi_x;
        U.rfac = P.radius / (P.radius + 1.0);
        for (int j = 0; j < 6; ++j) U.dl[j] = (NF >= 6) ? P.delta_f[j] : 0.0;
        const double dk = (NF == 7) ? P.delta_f[6] : 0.0;
        U.K1 = U.c2 * dk; U.K2 = 0.5 * U.c2 * U.c2 * dk;
    }
    U.K4 = 0.5 * U.c2c * U.c2c;
    *Up = U;
}

__device__ __forceinline__ double mask_double(double v, bool keep)
{   // v or +0.0, as integer logic: never becomes a branch around the (expensive) producer of v
    return __longlong_as_double(__double_as_longlong(v) & (keep ? -1ll : 0ll));
}

// One residual block, no branches.  Measured on B200: the sweep is bound by instruction issue / operand
// delivery (about 1.9 cycles per warp instruction of ANY kind with two warps per sub-partition), not by
// latency, so the body is written for the fewest instructions: no per-pixel finiteness flags (a non-finite
// residual, Jacobian or depth step makes sum r^2, sum d^2 or sum |step|^2 non-finite -- all terms are
// squares -- and the controller tests those), one mask per reciprocal, shared products for the two
// projected Jacobian vectors.  The scalar terms are added to acc / S and the block's two projected
// vectors are returned -- unless the block needs the exact path (a real block whose e^Te left the fast
// range at x or at the evaluation point): then nothing is added, the vectors are zero and `true` is returned.
template <int NF, bool INIT>
__device__ __forceinline__ bool pixel_fast(const Loaded &L, bool inb, int index, const SweepU &U, bool e_role,
                                           double *__restrict__ d_cand, double (&acc)[TAcc<NF>::NS], SweepScalars &S,
                                           double (&kv)[NF > 0 ? NF : 1], double (&sv)[NF > 0 ? NF : 1], double &ks, double &ss)
{
    const double x = L.p.x, y = L.p.y, d = L.d;
    const double xy = x * y, xx1 = fma(x, x, 1.0), yy1 = fma(y, y, 1.0);
    double a0 = fma(-x, U.v2, U.v0), a1 = fma(-y, U.v2, U.v1);
    double b0 = fma(-xy, U.w0, fma(xx1, U.w1, -(y * U.w2)));
    double b1 = fma(-yy1, U.w0, fma(xy, U.w1, x * U.w2));
    double dc = d, mj0 = 0.0, mt0 = 0.0, mj1 = 0.0, mt1 = 0.0, stp = 0.0;
    bool fast = true;
    if (!INIT) {
        // ---- candidate step at x: delta_d = -q e^T (r + F delta_f)
        const double ak = fma(U.k, L.a.y, L.a.x), beta = U.c2 * ak;
        const double p0 = fma(d, a0, b0), p1 = fma(d, a1, b1);
        const double r0 = fma(-beta, p0, L.u.x), r1 = fma(-beta, p1, L.u.y);
        const double e0 = -(beta * a0), e1 = -(beta * a1);
        const double ee = fma(e0, e0, e1 * e1);
        fast = bits_in_range(ee, U.lo_x, U.hi_x);
        const double q = mask_double(sweep_rcp(ee) * U.rfac, fast);
        double m0 = 0.0, m1 = 0.0;
        if (NF >= 6) {
            // F delta_f = -beta (d A dv + B dw) - dbeta p dk; the increments of A v and B w are reused for the candidate
            const double da0 = fma(-x, U.dl[2], U.dl[0]), da1 = fma(-y, U.dl[2], U.dl[1]);
            const double db0 = fma(-xy, U.dl[3], fma(xx1, U.dl[4], -(y * U.dl[5])));
            const double db1 = fma(-yy1, U.dl[3], fma(xy, U.dl[4], x * U.dl[5]));
            m0 = -(beta * fma(d, da0, db0));
            m1 = -(beta * fma(d, da1, db1));
            if (NF == 7) {
                const double dbk = fma(-ak, U.K2, U.K1 * L.a.y);          // (dbeta/dk) dk
                m0 = fma(-dbk, p0, m0);
                m1 = fma(-dbk, p1, m1);
            }
            a0 += da0; a1 += da1; b0 += db0; b1 += db1;
        }
        const double delta_e = -(q * fma(e0, r0 + m0, e1 * (r1 + m1)));
        const double j0 = fma(e0, delta_e, m0), j1 = fma(e1, delta_e, m1);                    // J delta
        dc = d + delta_e;
        if (inb) d_cand[index] = dc;
        mj0 = j0; mt0 = fma(0.5, j0, r0); mj1 = j1; mt1 = fma(0.5, j1, r1);    // model cost term (J delta)^T (r + J delta / 2)
        stp = delta_e;                                                       // |step|^2 term of this depth
    }
    // ---- evaluation at the candidate (INIT: at the start point)
    const double akc = fma(U.kc, L.a.y, L.a.x), bc = U.c2c * akc;
    const double p0 = fma(dc, a0, b0), p1 = fma(dc, a1, b1);
    const double r0 = fma(-bc, p0, L.u.x), r1 = fma(-bc, p1, L.u.y);
    const double e0 = -(bc * a0), e1 = -(bc * a1);
    const double ee = fma(e0, e0, e1 * e1);
    fast = fast && bits_in_range(ee, U.lo_c, U.hi_c);
    const bool slow = inb && !fast;
    // scalar terms (a padded record contributes zeros; a block on the exact path adds them there)
    if (!slow) {
        acc[0] = fma(r0, r0, fma(r1, r1, acc[0]));
        acc[1] = fma(dc, dc, acc[1]);
        if (!INIT) { acc[2] = fma(mj0, mt0, fma(mj1, mt1, acc[2])); acc[3] = fma(stp, stp, acc[3]); }
        S.gmax = umax64(S.gmax, dbits(fma(e0, r0, e1 * r1)) & 0x7fffffffffffffffull);     // |e^T r|
        if (INIT) S.eemax = umax64(S.eemax, dbits(ee));                // (only the first evaluation's max e^Te is used: lm_controller.h on_eval, phase 0)
    }
    if (NF > 0) {
        const double mu = mask_double(sweep_rsqrt(ee), fast);         // 1/|e|
        const double c = mu * e0, s = mu * e1;                        // unit depth-column direction
        const double mc = e_role ? c : -s, ms = e_role ? s : c;       // the direction this lane keeps: e or n = (-s, c)
        // F^T (mc, ms) and F^T (-ms, mc) share their products (F = -beta [d A | B | (dbeta/beta) p])
        const double P0 = bc * mc, P1 = bc * ms;
        const double nd = -dc;
        const double t1 = fma(x, P0, y * P1), t2 = fma(y, P0, -(x * P1));
        kv[0] = nd * P0;         sv[0] = dc * P1;
        kv[1] = nd * P1;         sv[1] = kv[0];
        kv[2] = dc * t1;         sv[2] = dc * t2;
        kv[3] = fma(y, t1, P1);  sv[3] = fma(y, t2, P0);
        kv[4] = fma(-x, t1, -P0); sv[4] = fma(-x, t2, P1);
        kv[5] = t2;              sv[5] = -t1;
        if (NF == 7) {
            const double ndbc = fma(akc, U.K4, -(U.c2c * L.a.y));     // -(dbeta/dk) at the evaluation point
            kv[6] = ndbc * fma(p0, mc, p1 * ms);
            sv[6] = ndbc * fma(p1, mc, -(p0 * ms));
        }
        ks = fma(mc, r0, ms * r1);
        ss = fma(mc, r1, -(ms * r0));
    } else {
        ks = 0.0; ss = 0.0;
    }
    return slow;
}

// rank-1 updates of one residual block pair: kv (own pixel, own direction) and pv (partner's pixel, own direction)
template <int NF>
__device__ __forceinline__ void rank1_update(const double (&kv)[NF > 0 ? NF : 1], const double (&pv)[NF > 0 ? NF : 1], double ks,
                                             double ps, double (&acc)[TAcc<NF>::NS])
{
    using T = TAcc<NF>;
    int t = 0;
#pragma unroll
    for (int j = 0; j < NF; ++j) {
        acc[T::oH + j] = fma(kv[j], ks, fma(pv[j], ps, acc[T::oH + j]));
#pragma unroll
        for (int c = j; c < NF; ++c, ++t) acc[T::oK + t] = fma(kv[j], kv[c], fma(pv[j], pv[c], acc[T::oK + t]));
    }
}

// ------------------------------------------------------------------------------------------
// CTA reduction.  Halving butterfly over the lanes of equal parity (xor 16, 8, 4, 2): at every level a
// lane keeps one half of its remaining values and sends the other half, so N values cost about N
// 64-bit shuffles instead of 4N.  Afterwards lane l holds the sums (over its 16 same-parity lanes) of the
// values with logical index  j + n4 b1 + n3 b2 + n2 b3 + n1 b4  (b_i = bit i of l), which go to
// part[warp][parity * N + index]; the warps are combined in warp order.  Fixed order => reproducible.
// ------------------------------------------------------------------------------------------
template <int NIN>
__device__ __forceinline__ void bfly_level(double *v, bool hi, int o)
{
    constexpr int NOUT = (NIN + 1) / 2;
#pragma unroll
    for (int j = 0; j < NOUT; ++j) {
        const double lo_v = v[j];
        const double hi_v = (j + NOUT < NIN) ? v[(j + NOUT < NIN) ? j + NOUT : 0] : 0.0;
        const double send = hi ? lo_v : hi_v, keep = hi ? hi_v : lo_v;
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
}

template <int N>
__device__ __forceinline__ void cta_reduce_sweep(double (&v)[N], const SweepScalars &S, double (*part)[kRowLd], double *row)
{
    constexpr int n1 = (N + 1) / 2, n2 = (n1 + 1) / 2, n3 = (n2 + 1) / 2, n4 = (n3 + 1) / 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b4 = (lane >> 4) & 1, b3 = (lane >> 3) & 1, b2 = (lane >> 2) & 1, b1 = (lane >> 1) & 1, b0 = lane & 1;
    bfly_level<N>(v, b4 != 0, 16);
    bfly_level<n1>(v, b3 != 0, 8);
    bfly_level<n2>(v, b2 != 0, 4);
    bfly_level<n3>(v, b1 != 0, 2);
    {
        const int s1 = n4 * b1, s2 = s1 + n3 * b2, s3 = s2 + n2 * b3, s4 = s3 + n1 * b4;
#pragma unroll
        for (int j = 0; j < n4; ++j) {
            const bool ok = (j + s1 < n3) && (j + s2 < n2) && (j + s3 < n1) && (j + s4 < N);
            if (ok) part[warp][b0 * N + j + s4] = v[j];
        }
    }
    // integer words: max |e^T r|, max e^Te (bit patterns of non-negative doubles), flags
    {
        const unsigned gh = __reduce_max_sync(0xffffffffu, (unsigned)(S.gmax >> 32));
        const unsigned gl = __reduce_max_sync(0xffffffffu, ((unsigned)(S.gmax >> 32) == gh) ? (unsigned)S.gmax : 0u);
        const unsigned eh = __reduce_max_sync(0xffffffffu, (unsigned)(S.eemax >> 32));
        const unsigned el = __reduce_max_sync(0xffffffffu, ((unsigned)(S.eemax >> 32) == eh) ? (unsigned)S.eemax : 0u);
        const unsigned fl = __reduce_or_sync(0xffffffffu, S.flags);
        if (lane == 0) {
            part[warp][2 * N + 0] = __longlong_as_double((long long)(((unsigned long long)gh << 32) | gl));
            part[warp][2 * N + 1] = __longlong_as_double((long long)(((unsigned long long)eh << 32) | el));
            part[warp][2 * N + 2] = __longlong_as_double((long long)(unsigned long long)fl);
        }
    }
    __syncthreads();
    if (tid < 2 * N) {
        double x = part[0][tid];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) x += part[w][tid];
        row[tid] = x;
    } else if (tid < 2 * N + 3) {
        unsigned long long x = dbits(part[0][tid]);
        if (tid == 2 * N + 2) { for (int w = 1; w < kWarps; ++w) x |= dbits(part[w][tid]); }
        else                  { for (int w = 1; w < kWarps; ++w) x = umax64(x, dbits(part[w][tid])); }
        row[tid] = __longlong_as_double((long long)x);
    }
    __syncthreads();
}

// plain variant (all values reduced over all lanes): used for the exception sums (rare)
template <int NS>
__device__ __forceinline__ void cta_reduce_sums(const double (&v)[NS], double (*wpart)[kRowLd], double *row)
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

// ------------------------------------------------------------------------------------------
// Warp-parallel controller steps (one warp, shared-memory state).  Same arithmetic as
// LmController::on_eval_stored / solve_step (which the host-stepped RANSAC solver keeps using).
// ------------------------------------------------------------------------------------------
template <int NF>
__device__ __forceinline__ int ctl_on_eval(LmController &c)
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
    for (int o = 4; o > 0; o >>= 1) { g = fmax(g, __shfl_xor_sync(0xffffffffu, g, o)); xs += __shfl_xor_sync(0xffffffffu, xs, o); }
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
__device__ __forceinline__ int ctl_solve(LmController &c, const ExcSums *exc, double (*Lm)[8])
{
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

// Next phase parameters from the controller state (one thread).
template <int NF>
__device__ __forceinline__ void publish_phase(const LmController &c, PhaseParams &P, SweepU &U, int nx, bool which_changed)
{
    if (which_changed) P.which_x ^= 1;                           // the candidate became x
    Motion mo = P.base, ca = P.base;
    if (NF >= 6) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            mo.v[j] = c.f[j]; mo.w[j] = c.f[3 + j];
            ca.v[j] = c.f[j] + c.delta_f[j]; ca.w[j] = c.f[3 + j] + c.delta_f[3 + j];
        }
    }
    if (NF == 7) { mo.k = c.f[6]; ca.k = c.f[6] + c.delta_f[6]; }
    if (nx == (int)LM_DONE && c.termination == RSDSFM_FAILURE) mo = P.base;   // Ceres restores the start values
    P.mot = mo; P.cand = ca;
#pragma unroll
    for (int j = 0; j < kMaxNF; ++j) P.delta_f[j] = c.delta_f[j];
    P.radius = c.radius;
    P.ee_fast_min = c.ee_fast_min;
    P.first = 0;
    P.next = nx;
    sweep_uniforms<NF>(&P, &U);
}

// The whole controller step of a phase, run by ONE warp as one contiguous, mostly straight-line piece of
// code (it executes once per phase with a cold instruction cache -- the sweep evicts it -- so taken branches
// and calls, not arithmetic, are what it costs): judge the candidate (LmController::on_candidate), take
// over the evaluation sums, the Ceres bookkeeping of the new point, the damped Cholesky solve, and the next
// phase's parameters.  fin: the combined row.  Returns (and leaves in s_flag[6]) 1 when the current point
// has listed (clamped) pixels: their sums need the whole CTA, and the caller finishes the step.
template <int NF>
__device__ __noinline__ int controller_warp(LmController &c, PhaseParams &P, SweepU &U, const double *fin, int *s_flag,
                                            const unsigned int *s_ne, const unsigned int *s_ne_all, int slot_cur, int slot_spec,
                                            bool run_init, unsigned int exc_cap, double (*Lm)[8])
{
    using T = TAcc<NF>;
    using RW = Row<NF>;
    constexpr int NS = T::NS;
    const int lane = threadIdx.x & 31;
    // flags of the exact path, plus what the sums themselves say: the sweep keeps no per-pixel finiteness flags,
    // a non-finite residual / depth step / Jacobian entry shows up as a non-finite sum (1: residual, 2: evaluation, 4: step)
    unsigned int flags = 0u;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int j = lane + 32 * q;
        if (j < 2 * NS && not_finite(fin[j < 2 * NS ? j : 0])) {
            const int jj = j >= NS ? j - NS : j;
            flags |= (jj == 0) ? 3u : (jj == 1) ? 6u : (jj == 2) ? 0u : (jj == 3) ? 4u : 2u;
        }
    }
    flags = __reduce_or_sync(0xffffffffu, flags) | (unsigned int)dbits(fin[RW::oFLAGS]);
    if (lane == 0) s_flag[7] = (int)flags;
    int nx = (int)LM_RUN_A;
    if (!run_init) {
        if (lane == 0) {
            CandSums cs;
            cs.mcc = fin[2] + fin[NS + 2]; cs.step_sq = fin[3] + fin[NS + 3]; cs.cand_cost = 0.5 * (fin[0] + fin[NS]);
            cs.bad_step = (flags & 4u) ? 1.0 : 0.0; cs.bad_cand = (flags & 1u) ? 1.0 : 0.0;
            nx = (int)c.on_candidate(cs);
        }
        nx = __shfl_sync(0xffffffffu, nx, 0);
    }
    const bool accepted = (nx == (int)LM_RUN_A);
    const int cur = run_init ? slot_cur : (accepted ? slot_spec : slot_cur);
    unsigned int ne = s_ne[cur];                                 // this GPU's list
    if (ne > exc_cap) ne = exc_cap;
    const int need_cta = (NF > 0 && s_ne_all[cur] > 0u) ? 1 : 0;   // listed pixels on ANY GPU of a row split
    if (lane == 0) { s_flag[1] = nx; s_flag[3] = accepted ? 1 : 0; s_flag[4] = cur; s_flag[2] = (int)ne; s_flag[6] = need_cta; }
    if (need_cta) return 1;
    if (accepted) {
        // the evaluation sums of this pass describe the (new) current point: EvalSums in place
        if (lane < kTri) {
            c.ev.G1[lane] = (lane < T::TRI) ? fin[T::oK + (lane < T::TRI ? lane : 0)] : 0.0;
            c.ev.G2[lane] = (lane < T::TRI) ? fin[NS + T::oK + (lane < T::TRI ? lane : 0)] : 0.0;
        }
        if (lane < kMaxNF) {
            c.ev.h1[lane] = (lane < NF) ? fin[T::oH + (lane < NF ? lane : 0)] : 0.0;
            c.ev.h2[lane] = (lane < NF) ? fin[NS + T::oH + (lane < NF ? lane : 0)] : 0.0;
        }
        if (lane == 31) {
            c.ev.cost = 0.5 * (fin[0] + fin[NS]); c.ev.sumsq_d = fin[1] + fin[NS + 1];
            c.ev.gmax_e = fin[RW::oGMAX];
            c.ev.bad = (flags & 2u) ? 1.0 : 0.0; c.ev.ee_max = fin[RW::oEEMAX];
        }
        __syncwarp();
        nx = ctl_on_eval<NF>(c);
    }
    while (nx == (int)LM_SOLVE) nx = ctl_solve<NF>(c, nullptr, Lm);
    __syncwarp();
    if (lane == 0) publish_phase<NF>(c, P, U, nx, !run_init && accepted);
    return 0;
}

// out-of-line copies for the (rare) CTA-wide path with listed pixels
template <int NF> __device__ __noinline__ int ctl_on_eval_cold(LmController &c) { return ctl_on_eval<NF>(c); }
template <int NF> __device__ __noinline__ int ctl_solve_cold(LmController &c, const ExcSums *exc, double (*Lm)[8]) { return ctl_solve<NF>(c, exc, Lm); }
template <int NF> __device__ __noinline__ void publish_phase_cold(const LmController &c, PhaseParams &P, SweepU &U, int nx, bool which_changed)
{
    publish_phase<NF>(c, P, U, nx, which_changed);
}

constexpr unsigned long long kWatchdogNs = 4000000000ull;   // 4 s: a stuck grid barrier aborts the solve
constexpr int kExcSlots = 3;

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
__device__ __noinline__ bool sort_exceptions(const ExcEntry *list, int ne, unsigned long long *keys, int cap)
{
    const int tid = threadIdx.x;
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

// Listed pixels (rare; whole CTA).  mode 0: their radius-independent part F^TF, F^Tr (joins G1, h1 of a new
// point); mode 1: the radius-dependent correction sum q fe fe^T, sum q fe (e^T r) at radius R.  out[kExcVals].
template <int NF>
__device__ __noinline__ void exc_sums(int mode, const ExcEntry *cur_exc, int ne, const unsigned long long *xkeys, bool xsorted,
                                      double R, double lo, double hi, double (*part)[kRowLd], double *out)
{
    constexpr int NFa = NF > 0 ? NF : 1;
    const int tid = threadIdx.x;
    double a[kExcVals];
#pragma unroll
    for (int j = 0; j < kExcVals; ++j) a[j] = 0.0;
    for (int k = tid; k < ne; k += kThreads) {
        const int slot = xsorted ? (int)(unsigned int)(xkeys[k] & 0xffffffffull) : k;
        ExcEntry X;
        for (int w = 0; w < (int)(sizeof(ExcEntry) / sizeof(double)); ++w)
            reinterpret_cast<double *>(&X)[w] = __ldcg(reinterpret_cast<const double *>(cur_exc + slot) + w);
        if (mode == 0) {
            int t = 0;
#pragma unroll
            for (int j = 0; j < NF; ++j) {
                a[kTri + j] += fma(X.F0[j], X.r0, X.F1[j] * X.r1);
#pragma unroll
                for (int c = j; c < NF; ++c, ++t) a[t] += fma(X.F0[j], X.F0[c], X.F1[j] * X.F1[c]);
            }
        } else {
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
    }
    cta_reduce_sums<kExcVals>(a, part, out);
}

// All-reduce of vals[0..n) over the GPUs of a row split (whole CTA; vals in shared memory, identical in all CTAs of
// a GPU).  Columns [0, n_sum) and [tail, n) are added in GPU order, column i_or is or-ed, the others (bit patterns of
// non-negative doubles) are max-ed.  seq: number of this exchange within the solve (the same on every GPU).  false: watchdog.
__device__ __noinline__ bool peer_allreduce(const PeerInfo &pi, double *vals, int n, int n_sum, int i_or, int tail, unsigned int seq,
                                            const int *abort_flag)
{
    const int tid = threadIdx.x;
    const unsigned long long tag = ((unsigned long long)pi.epoch << 32) | (unsigned long long)(seq + 1u);
    const size_t slot = (size_t)(seq % (unsigned)kMailSlots) * kMaxPeers * kMailLd;
    if (blockIdx.x == 0 && tid < n)
        for (int g = 0; g < pi.n; ++g) st_tagged(pi.mail[g] + slot + (size_t)pi.me * kMailLd + tid, vals[tid], tag);
    int err = 0;
    if (tid < n) {
        const Tagged *box = pi.mail[pi.me] + slot + tid;
        double x = 0.0;
        unsigned long long xb = 0ull;
        const unsigned long long t0 = globaltimer();
        for (int g = 0; g < pi.n; ++g) {
            Tagged t = ld_tagged(box + (size_t)g * kMailLd);
            while (t.tag != tag) {
                __nanosleep(40);
                t = ld_tagged(box + (size_t)g * kMailLd);
                if (t.tag != tag && (globaltimer() - t0 > 4000000000ull || __ldcg(abort_flag))) { err = 1; break; }
            }
            if (err) break;
            if (tid < n_sum || tid >= tail) x += t.v;
            else if (tid == i_or) xb |= (unsigned long long)__double_as_longlong(t.v);
            else { const unsigned long long b = (unsigned long long)__double_as_longlong(t.v); xb = b > xb ? b : xb; }
        }
        vals[tid] = (tid < n_sum || tid >= tail) ? x : __longlong_as_double((long long)xb);
    }
    return __syncthreads_or(err) == 0;
}

struct SolveArgs {
    RefineData D;
    double *d0, *d1;
    LmShared *sh;
    double *partials;            // [2][gridDim.x][Row<NF>::NV]: rows of even / odd phases
    ExcEntry *exc;               // [kExcSlots][exc_cap]
    unsigned int exc_cap;
    const double *z_in;
    int z_stride;
    double *out;
    int invert_out;
    double *zstats;
    PeerInfo peers;              // n = 1: a solve on one GPU
};

// Ring bookkeeping: tile "uses" are numbered from kernel start; use u lives in stage u % kStages and completes
// phase (u / kStages) of that stage's full barrier.
struct RingPos {
    int s;
    unsigned par;
    __device__ __forceinline__ void advance() { if (++s == kStages) { s = 0; par ^= 1u; } }
};

// ------------------------------------------------------------------------------------------
// One sweep over this CTA's tiles: one residual block per thread and step.
//  * Ring refill by rotation: at step k the leader of warp k % 8 refills the stage that held tile
//    k - kRefillLag (waiting, if it must, until every warp has released it) with tile k - kRefillLag + kStages.
//    Every tile is queued by a warp that is known in advance: no producer state, and the cost of queueing
//    (two bulk copies per tile) is spread over all warps instead of making warp 0 the straggler.
//  * The full barrier of the NEXT tile is tested at the top of a step; the blocking wait is only entered
//    when that early test failed.
//  (Variants that were measured and dropped: rank-1 updates software-pipelined one step behind the chain,
//   with and without artificial dependences that spread them over the chain's latencies -- both slower than
//   this plain order, the sweep is issue-bound; two residual blocks per step; a producer warp.)
// ------------------------------------------------------------------------------------------
constexpr int kRefillLag = 5;
// The residual blocks are dealt to kStrips strips (tile t belongs to strip t % kStrips), every strip is summed by one CTA
// in tile order into a row of its own, and the rows are combined in strip order: the sums -- and with them the whole
// solve -- do not depend on the grid.  A full-GPU solve has one strip per CTA; a solve that shares the GPU with others
// (rsdsfm_refine_rectify_sequence: grid = a fraction of the SMs) walks through several strips per phase, CTA b taking
// strips b, b + grid, ...  Bit-identical results either way.  (The strip count is the B200's SM count and part of the
// result's definition: grids that do not divide it -- a part with fewer SMs -- still work, with uneven shares.)
constexpr int kStrips = kNumSMsB200;

template <int NF, bool INIT>
__device__ __forceinline__ void sweep(const SolveArgs &A_, const PhaseParams &P, const SweepU &Us, const uint32_t *A_s, const int *vs, int strip, int n_my,
                                      RingPos &cons, unsigned int &consumed, const double *dx,
                                      double *dcand, int elist, double (&acc)[TAcc<NF>::NS], SweepScalars &S)
{
    // strip `strip` of this CTA = strip v of the solve (see kStrips): tile uses vs[strip] .. vs[strip + 1] - 1 of this
    // CTA's n_my per phase, tiles v, v + kStrips, ...
    const int k0 = vs[strip], k1 = vs[strip + 1];
    int idx = ((int)blockIdx.x + strip * (int)gridDim.x) * kTile + (int)threadIdx.x;   // this thread's residual block of the step
    constexpr int NFa = NF > 0 ? NF : 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    const RefineData D = A_.D;
    const bool e_role = (tid & 1) != 0;
    unsigned int *n_exc = &A_.sh->n_exc[elist];
    ExcEntry *elist_p = A_.exc + (size_t)elist * A_.exc_cap;

    const SweepU U = Us;                   // registers
    RingPos rc = cons;                     // consumer position (registers)
    // shared-window addresses, computed once: ring, this thread's record inside a stage, the barriers
    // (read back from shared memory: values the compiler can recompute from %cluster_ctarank it DOES recompute, every step)
    const uint32_t ring_s = ((const volatile uint32_t *)A_s)[0], full_s = ((const volatile uint32_t *)A_s)[1], empty_s = ((const volatile uint32_t *)A_s)[2];
    const uint32_t rec16 = (uint32_t)tid * 16u, rec8 = (uint32_t)(3 * kTile * 16) + (uint32_t)tid * 8u;
#ifdef LM_DBG_WAITCLK
    long long dbg_wait = 0; const long long dbg_t0 = clock64();
#endif
    bool ready = false;                    // the early test of this step's full barrier succeeded
    int kduty = k0 + ((warp - k0) & (kWarps - 1));   // this warp's next refill duty: the steps k with k % 8 == warp
#pragma unroll 1
    for (int k = k0; k < k1; ++k, idx += kStrips * kTile) {
        const bool inb = idx < D.m;
#ifdef LM_DBG_WAITCLK
        const long long w0_ = clock64();
#endif
#ifdef LM_DBG_NOLOAD
        if (consumed + (unsigned)(k - k0) < (unsigned)kStages) mbar_wait_s(full_s + 8u * (uint32_t)rc.s, rc.par);
#else
        if (!ready) mbar_wait_s(full_s + 8u * (uint32_t)rc.s, rc.par);
#endif
#ifdef LM_DBG_WAITCLK
        dbg_wait += clock64() - w0_;
#endif
        Loaded L;
        {
            const uint32_t st = ring_s + (uint32_t)rc.s * (uint32_t)sizeof(Stage);
            L.p = lds_f64x2(st + rec16); L.u = lds_f64x2(st + (uint32_t)(kTile * 16) + rec16);
            L.a = lds_f64x2(st + (uint32_t)(2 * kTile * 16) + rec16); L.d = lds_f64(st + rec8);
        }
        const int s_now = rc.s;
        const unsigned par_now = rc.par;
        rc.advance();
#ifndef LM_DBG_NOLOAD
        ready = mbar_test_s(full_s + 8u * (uint32_t)rc.s, rc.par);        // next tile: result needed only at the top of the next step
#endif
        // ---- the residual block: straight-line
        double kv[NFa], sv[NFa], pv[NFa], ks, ss, ps = 0.0;
#ifdef LM_DBG_NOCOMPUTE
        acc[0] += L.p.x + L.p.y + L.u.x + L.u.y + L.a.x + L.a.y + L.d;
        if (!INIT && inb) dcand[idx] = L.d;
        ks = 0.0; ss = 0.0;
#pragma unroll
        for (int j = 0; j < NFa; ++j) { kv[j] = 0.0; sv[j] = 0.0; }
        const bool slow = false;
#else
        const bool slow = pixel_fast<NF, INIT>(L, inb, idx, U, e_role, dcand, acc, S, kv, sv, ks, ss);
#endif
        if (__any_sync(0xffffffffu, slow)) {
            // rare: some block of this warp needs the exact treatment (it added nothing above, its vectors are zero):
            // its lane replaces them before the partners exchange
            if (slow) {
                const Loaded Lc = L;
                PixOut<NF> X;
                pixel_exact<NF, INIT>(&Lc, idx, &P, dcand, n_exc, &A_.sh->exc_overflow, elist_p, A_.exc_cap, &X);
                acc[0] += X.rr; acc[1] += X.dd2;
                if (!INIT) { acc[2] += X.mcc; acc[3] += X.stp; }
                S.gmax = umax64(S.gmax, X.gb); S.eemax = umax64(S.eemax, X.eb); S.flags |= X.flags;
#pragma unroll
                for (int j = 0; j < NFa; ++j) { kv[j] = X.kv[j]; sv[j] = X.sv[j]; }
                ks = X.ks; ss = X.ss;
            }
        }
        if (NF > 0) {
#pragma unroll
            for (int j = 0; j < NF; ++j) pv[j] = __shfl_xor_sync(0xffffffffu, sv[j], 1);
            ps = __shfl_xor_sync(0xffffffffu, ss, 1);
            rank1_update<NF>(kv, pv, ks, ps, acc);
        }
        // ---- this warp is done with the stage: release it (one arrival per warp; the warp is converged here and
        // every lane's loads of the tile have long been consumed)
#ifndef LM_DBG_NOLOAD
        if (lane == 0) mbar_arrive_s(empty_s + 8u * (uint32_t)s_now);
#endif
        // ---- refill duty of this step (warp-uniform test): the stage that held tile k - kRefillLag gets tile k - kRefillLag + kStages
        const int kd = k - kRefillLag;
#ifdef LM_DBG_NOLOAD
        if (false) {
#else
        if (k == kduty) {
            kduty += kWarps;
#endif
          if ((kd >= 0) & (kd + kStages < n_my)) {
            if (lane == 0) {
                int sd = s_now - kRefillLag;                       // ring position of tile kd, from this step's
                unsigned pd = par_now;
                if (sd < 0) { sd += kStages; pd ^= 1u; }
                mbar_wait_s(empty_s + 8u * (uint32_t)sd, pd);     // every warp has released tile kd
                int tile = (idx - tid) / kTile + (kStages - kRefillLag) * kStrips;   // tile use kd + kStages, if it is in this strip
                if (kd + kStages >= k1) {                          // no: a later strip of this CTA
                    int rs = 0;
                    while (kd + kStages >= vs[rs + 1]) ++rs;
                    tile = (int)blockIdx.x + rs * G + (kd + kStages - vs[rs]) * kStrips;
                }
                const uint32_t st = ring_s + (uint32_t)sd * (uint32_t)sizeof(Stage), fb = full_s + 8u * (uint32_t)sd;
                mbar_expect_tx_s(fb, (unsigned)sizeof(Stage));
                bulk_g2s_s(st, D.blk + (size_t)tile * (3 * kTile), (unsigned)(3 * kTile * sizeof(double2)), fb);
                bulk_g2s_s(st + (uint32_t)(3 * kTile * 16), dx + (size_t)tile * kTile, (unsigned)(kTile * sizeof(double)), fb);
            }
          }
        }
    }
#ifdef LM_DBG_WAITCLK
    if (!INIT && lane == 0 && (blockIdx.x == 0 || blockIdx.x == 77) && consumed > 20u * (unsigned)n_my && consumed < 21u * (unsigned)n_my + 20u)
        printf("cta %d warp %d: sweep %lld cycles, waiting for tiles %lld cycles, %d steps\n", (int)blockIdx.x, tid >> 5, clock64() - dbg_t0, dbg_wait, n_my);
#endif
    consumed += (unsigned)(k1 - k0);
    cons = rc;
}

// elected thread: queue the first `pre` tiles of the phase that starts at ring position `at` (= the consumer's
// position: everything before it has been consumed) and reads `dsrc`; `first_use` = number of the first tile use
__device__ __noinline__ void queue_phase_head(const RefineData D, const double *dsrc, Stage *stages, uint64_t *full, uint64_t *empty,
                                              RingPos at, unsigned int first_use, int pre, const int *vs)
{
    const int G = gridDim.x;
    int rs = 0;
#ifdef LM_DBG_NOLOAD
    if (first_use != 0u) return;
#endif
    fence_proxy_async();          // the ring may have served as scratch (sort keys) since its last tile
    for (int i = 0; i < pre; ++i) {
        if (first_use + (unsigned)i >= (unsigned)kStages) mbar_wait(&empty[at.s], at.par ^ 1u);   // the stage's previous use was released
        while (i >= vs[rs + 1]) ++rs;
        issue_tile(D, dsrc, (int)blockIdx.x + rs * G + (i - vs[rs]) * kStrips, &stages[at.s], &full[at.s]);
        at.advance();
    }
}

// whole CTA: wait for the `pre` prefetched tiles of a phase that will not run (or whose depth buffer was guessed
// wrong) and hand their stages back
__device__ __noinline__ void drain_prefetch(uint64_t *full, uint64_t *empty, RingPos *cons, unsigned int *consumed, int pre, bool release)
{
    const int lane = threadIdx.x & 31;
#ifdef LM_DBG_NOLOAD
    return;
#endif
    for (int u = 0; u < pre; ++u) {
        mbar_wait(&full[cons->s], cons->par);
        if (release) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[cons->s]);
        }
        cons->advance();
    }
    *consumed += (unsigned)pre;
}

template <int NF>
__global__ void __launch_bounds__(kThreads, 1) k_lm_solve(const SolveArgs A_)
{
    using T = TAcc<NF>;
    using RW = Row<NF>;
    constexpr int NS = T::NS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Stage *stages = reinterpret_cast<Stage *>(smem_raw);
    __shared__ __align__(8) uint64_t full[kStages], empty[kStages];
    __shared__ PhaseParams P;
    __shared__ SweepU U;
    __shared__ LmController s_ctl;
    __shared__ double fin[kRowLd];
    __shared__ double part[kWarps][kRowLd];
    __shared__ ExcSums s_exc;
    __shared__ double s_L[7][8];
    __shared__ int s_flag[8];     // [1] next, [2] n_exc of the current list, [3] accepted, [4] current slot, [5] error, [6] listed pixels: CTA-wide path, [7] flags
    __shared__ unsigned int s_ne[kExcSlots], s_ne_all[kExcSlots];   // listed pixels: on this GPU / on all GPUs of a row split
    __shared__ uint32_t s_addr[4];  // shared-window addresses of the ring and of the full / empty barriers

    const RefineData D = A_.D;
    double *const d0 = A_.d0, *const d1 = A_.d1;
    LmShared *const sh = A_.sh;
    ExcEntry *const exc = A_.exc;
    const unsigned int exc_cap = A_.exc_cap;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    const int NT = (D.m + kTile - 1) / kTile;
    const int n_rows = NT < kStrips ? NT : kStrips;               // strips that have tiles = rows of the exchange
    const int n_strips = ((int)blockIdx.x < n_rows) ? (n_rows - 1 - (int)blockIdx.x) / G + 1 : 0;   // strips of this CTA
    __shared__ int s_vs[kStrips + 2];                             // s_vs[i]: tile uses of this CTA (per phase) before its strip i
    if (tid == 0) {
        int at = 0;
        for (int i = 0; i < n_strips; ++i) { s_vs[i] = at; at += (NT - 1 - ((int)blockIdx.x + i * G)) / kStrips + 1; }
        s_vs[n_strips] = at;
        s_vs[n_strips + 1] = 0x7fffffff;                          // (stops the searches for a tile use's strip)
    }
    __syncthreads();
    const int n_my = s_vs[n_strips];
    const int pre = n_my < kStages ? n_my : kStages;              // tiles queued ahead of a phase
    unsigned int gen = 0;
    unsigned int xseq = 0;                                        // exchanges with the other GPUs of a row split so far
    unsigned int consumed = 0;                                    // tile uses consumed by this CTA since kernel start
    RingPos cons{0, 0u};                                          // every use before this position has been consumed; a phase's first `pre` tiles are queued ahead of it
    // exception lists: cur = list of the current point, spec = list the FUSED evaluation appends to,
    // zero = list that thread 0 of CTA 0 clears during this phase (it becomes `spec` of the next phase)
    int slot_cur = 0, slot_spec = 1, slot_zero = 2;

    if (blockIdx.x == 0 && tid == 0) sh->t_abs[0] = globaltimer();
    if (tid == 0) {
        s_addr[0] = smem_u32(stages); s_addr[1] = smem_u32(full); s_addr[2] = smem_u32(empty);
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
        const double lo = __ldcg(&sh->ctl.opt.min_lm_diagonal), hi = __ldcg(&sh->ctl.opt.max_lm_diagonal);
        P.min_diag = lo; P.max_diag = hi;
        // first evaluation: the Jacobi scale of a depth column is 1/(1+|e|) of that very point, so its scaled
        // diagonal e^Te/(1+|e|)^2 is monotone in e^Te: it lies inside [lo, hi] iff e^Te lies inside
        // [lo/(1-sqrt(lo))^2, hi/(1-sqrt(hi))^2].  A relative margin keeps every borderline block on the exact path.
        const double sl = sqrt(lo), shi = sqrt(hi);
        P.ee_first_min = (lo > 0.0 && sl < 1.0) ? lo / ((1.0 - sl) * (1.0 - sl)) * (1.0 + 1e-9) : ((lo > 0.0) ? INFINITY : 0.0);
        P.ee_first_max = (shi < 1.0) ? hi / ((1.0 - shi) * (1.0 - shi)) * (1.0 - 1e-9) : 1.7976931348623157e308;
        P.error = 0;
    }
    __syncthreads();
    if (tid == 32) sweep_uniforms<NF>(&P, &U);

    if (tid == 0 && P.next != LM_DONE) queue_phase_head(D, P.which_x ? d1 : d0, stages, full, empty, cons, consumed, pre, s_vs);
    __syncthreads();

    if (blockIdx.x == 0 && tid == 0) sh->t_abs[2] = globaltimer();
    for (;;) {
        if (P.next == LM_DONE || P.error) break;
        const bool run_init = (P.next == LM_RUN_A);
        double *dx = P.which_x ? d1 : d0;
        double *dcand = P.which_x ? d0 : d1;
        const unsigned long long t_begin = (blockIdx.x == 0 && tid == 0) ? globaltimer() : 0ull;
        if (blockIdx.x == 0 && tid == 0 && !run_init) sh->n_exc[slot_zero] = 0u;

        unsigned long long t_loop = 0ull;
#pragma unroll 1
        for (int strip = 0; strip < n_strips; ++strip) {
            double acc[NS];
#pragma unroll
            for (int j = 0; j < NS; ++j) acc[j] = 0.0;
            SweepScalars S;
            S.gmax = 0ull; S.eemax = 0ull; S.flags = 0u;
            if (run_init) sweep<NF, true>(A_, P, U, s_addr, s_vs, strip, n_my, cons, consumed, dx, dcand, slot_cur, acc, S);
            else          sweep<NF, false>(A_, P, U, s_addr, s_vs, strip, n_my, cons, consumed, dx, dcand, slot_spec, acc, S);
            // the candidate depths written above are read by TMA in the next phase: order them for the async proxy
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncthreads();
            if (t_begin) t_loop = globaltimer();                 // (with several strips: the earlier strips' reductions count as loop time)
            double *row = A_.partials + ((size_t)(gen & 1u) * kStrips + (size_t)((int)blockIdx.x + strip * G)) * RW::NV;
            cta_reduce_sweep<NS>(acc, S, part, row);
        }
        // ---- arrive; meanwhile another warp queues the next phase's first tiles (depth: from the buffer an ACCEPTED step makes current)
        if (tid == 32) queue_phase_head(D, run_init ? dx : dcand, stages, full, empty, cons, consumed, pre, s_vs);
        if (tid == 0) {
            __threadfence();
            atomicAdd(&sh->arrive, 1u);
            if (t_begin) {
                const unsigned long long t2 = globaltimer();
                if (t_loop == 0ull) t_loop = t2;                  // (a CTA without strips)
                atomicAdd(&sh->t_phase[run_init ? 4 : 7], t_loop - t_begin); atomicAdd(&sh->t_phase[run_init ? 5 : 8], t2 - t_loop);
            }
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
            drain_prefetch(full, empty, &cons, &consumed, pre, false);
            if (tid == 0) P.error = 1;
            __syncthreads();
            break;
        }
        const unsigned long long t_ctl = t_begin ? globaltimer() : 0ull;

        // ---- every CTA: combine the strips' rows in a fixed order (warp w: rows w, w+8, ...; lanes: columns)
        {
            constexpr int nv = RW::NV, ns2 = 2 * NS;
            const double *rows = A_.partials + (size_t)(gen & 1u) * kStrips * RW::NV;
            if (tid < kExcSlots) s_ne[tid] = __ldcg(&sh->n_exc[tid]);
            constexpr int kRowsPerWarp = (kStrips + kWarps - 1) / kWarps;      // 19
            double v[3] = {0.0, 0.0, 0.0};
            {
                double t[kRowsPerWarp][3];
#pragma unroll
                for (int u = 0; u < kRowsPerWarp; ++u) {
                    const int b = u * kWarps + warp;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int j = lane + 32 * c;
                        t[u][c] = (b < n_rows && j < nv) ? __ldcg(rows + (size_t)b * nv + j) : 0.0;
                    }
                }
#pragma unroll
                for (int u = 0; u < kRowsPerWarp; ++u)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int j = lane + 32 * c;
                        if (j < ns2) v[c] += t[u][c];
                        else if (j == RW::oFLAGS) v[c] = __longlong_as_double((long long)(dbits(v[c]) | dbits(t[u][c])));
                        else v[c] = __longlong_as_double((long long)umax64(dbits(v[c]), dbits(t[u][c])));
                    }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) { const int j = lane + 32 * c; if (j < nv) part[warp][j] = v[c]; }
            __syncthreads();
            if (tid < nv) {
                double x = part[0][tid];
                if (tid < ns2) { for (int w = 1; w < kWarps; ++w) x += part[w][tid]; }
                else if (tid == RW::oFLAGS) { unsigned long long b = dbits(x); for (int w = 1; w < kWarps; ++w) b |= dbits(part[w][tid]); x = __longlong_as_double((long long)b); }
                else { unsigned long long b = dbits(x); for (int w = 1; w < kWarps; ++w) b = umax64(b, dbits(part[w][tid])); x = __longlong_as_double((long long)b); }
                fin[tid] = x;
            }
            __syncthreads();
        }
        if (tid < kExcSlots) s_ne_all[tid] = s_ne[tid];
        if (A_.peers.n > 1) {
            // row split: the totals (and the numbers of listed pixels) of all GPUs, combined in GPU order
            if (tid < kExcSlots) fin[RW::NV + tid] = (double)s_ne[tid];
            __syncthreads();
            const bool ok = peer_allreduce(A_.peers, fin, RW::NV + kExcSlots, 2 * NS, RW::oFLAGS, RW::NV, xseq++, &sh->error);
            if (tid < kExcSlots) s_ne_all[tid] = (unsigned int)fin[RW::NV + tid];
            if (!ok) {
                if (tid == 0) { sh->error = 1; P.error = 1; }
                drain_prefetch(full, empty, &cons, &consumed, pre, false);
                __syncthreads();
                break;
            }
        }
        __syncthreads();
        const unsigned long long t_fin = t_begin ? globaltimer() : 0ull;
        // ---- the controller step: one warp, one contiguous piece of code (controller_warp)
        if (warp == 0)
            controller_warp<NF>(s_ctl, P, U, fin, s_flag, s_ne, s_ne_all, slot_cur, slot_spec, run_init, exc_cap, s_L);
        __syncthreads();
        const bool accepted = s_flag[3] != 0;
        if (!run_init) {
            // the list that is not current any more is cleared during the next phase and reused after it
            const int dead = accepted ? slot_cur : slot_spec;
            slot_cur = s_flag[4];
            slot_spec = slot_zero;
            slot_zero = dead;
        }
        bool drained = false;
        if constexpr (NF > 0) if (s_flag[6]) {
            // ---- listed pixels (rare): the whole CTA finishes the step.  Their sort keys live in the pixel ring,
            // so the prefetched tiles are given up first.
            if (accepted) {
                if (tid < kTri) {
                    s_ctl.ev.G1[tid] = (tid < T::TRI) ? fin[T::oK + (tid < T::TRI ? tid : 0)] : 0.0;
                    s_ctl.ev.G2[tid] = (tid < T::TRI) ? fin[NS + T::oK + (tid < T::TRI ? tid : 0)] : 0.0;
                } else if (tid >= 32 && tid < 32 + kMaxNF) {
                    const int j = tid - 32;
                    s_ctl.ev.h1[j] = (j < NF) ? fin[T::oH + (j < NF ? j : 0)] : 0.0;
                    s_ctl.ev.h2[j] = (j < NF) ? fin[NS + T::oH + (j < NF ? j : 0)] : 0.0;
                } else if (tid == 64) {
                    s_ctl.ev.cost = 0.5 * (fin[0] + fin[NS]); s_ctl.ev.sumsq_d = fin[1] + fin[NS + 1];
                    s_ctl.ev.gmax_e = fin[RW::oGMAX];
                    s_ctl.ev.bad = ((unsigned int)s_flag[7] & 2u) ? 1.0 : 0.0; s_ctl.ev.ee_max = fin[RW::oEEMAX];
                }
            }
            __syncthreads();
            unsigned long long *xkeys = reinterpret_cast<unsigned long long *>(smem_raw);
            drain_prefetch(full, empty, &cons, &consumed, pre, true);
            drained = true;
            __syncthreads();
            const bool xsorted = sort_exceptions(exc + (size_t)s_flag[4] * exc_cap, s_flag[2], xkeys, 16384);
            if (accepted) {
                exc_sums<NF>(0, exc + (size_t)s_flag[4] * exc_cap, s_flag[2], xkeys, xsorted, 0.0, 0.0, 0.0, part, fin);
                if (A_.peers.n > 1) peer_allreduce(A_.peers, fin, kExcVals, kExcVals, -1, kExcVals, xseq++, &sh->error);
                if (tid < T::TRI) s_ctl.ev.G1[tid] += fin[tid];
                if (tid < NF) s_ctl.ev.h1[tid] += fin[kTri + tid];
                __syncthreads();
                if (warp == 0) {
                    const int nx = ctl_on_eval_cold<NF>(s_ctl);
                    if (lane == 0) s_flag[1] = nx;
                }
                __syncthreads();
            }
            // (re)solve at the current radius; the clamped-pixel correction is summed by the whole CTA
            while (s_flag[1] == (int)LM_SOLVE) {
                exc_sums<NF>(1, exc + (size_t)s_flag[4] * exc_cap, s_flag[2], xkeys, xsorted, s_ctl.radius, s_ctl.opt.min_lm_diagonal,
                             s_ctl.opt.max_lm_diagonal, part, fin);
                if (A_.peers.n > 1) peer_allreduce(A_.peers, fin, kExcVals, kExcVals, -1, kExcVals, xseq++, &sh->error);
                if (tid < kTri) s_exc.S[tid] = fin[tid];
                if (tid < kMaxNF) s_exc.rhs[tid] = fin[kTri + tid];
                __syncthreads();
                if (warp == 0) {
                    const int nx = ctl_solve_cold<NF>(s_ctl, &s_exc, s_L);
                    if (lane == 0) s_flag[1] = nx;
                }
                __syncthreads();
            }
            if (tid == 0) publish_phase_cold<NF>(s_ctl, P, U, s_flag[1], !run_init && accepted);
            __syncthreads();
        }
        if (t_begin) {
            const unsigned long long t_end = globaltimer();
            atomicAdd(&sh->t_phase[run_init ? 10 : 11], t_end - t_fin);
            atomicAdd(&sh->t_phase[run_init ? 6 : 9], t_end - t_ctl);
            atomicAdd(&sh->t_phase[run_init ? 0 : 2], t_end - t_begin); atomicAdd(&sh->t_phase[run_init ? 1 : 3], 1ull);
        }
        gen++;
        // ---- the depth prefetch assumed "accepted" (or INIT): anything else reloads the first tiles
        const bool spec_ok = run_init || accepted;
        if (P.next != LM_DONE && (!spec_ok || drained)) {
            if (!drained) drain_prefetch(full, empty, &cons, &consumed, pre, true);
            __syncthreads();
            if (tid == 0) queue_phase_head(D, P.which_x ? d1 : d0, stages, full, empty, cons, consumed, pre, s_vs);
        } else if (P.next == LM_DONE && !drained) {
            drain_prefetch(full, empty, &cons, &consumed, pre, false);   // leave no bulk copy in flight when the CTA exits
        }
    }

    if (blockIdx.x == 0 && tid == 0) sh->t_abs[3] = globaltimer();
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
    // zstats (nullable): per-CTA rows {sum z, max z, max -z} of what was written, for the sign fix
    // and depth range of main.cc:466-489 -- saves the rectification stage a pass over z
    double zs[1] = {0.0}, zm[2] = {-INFINITY, -INFINITY};
    // (four loads in flight per thread: with one, 8 warps per SM leave this pass latency bound at ~0.4 TB/s;
    // the per-thread order of the sums is that of the plain strided loop)
    constexpr int kEpi = 4;
    const int stride = G * kThreads;
    for (int i0 = blockIdx.x * kThreads + tid; i0 < D.m; i0 += kEpi * stride) {
        double dv[kEpi];
#pragma unroll
        for (int u = 0; u < kEpi; ++u) {
            const int i = i0 + u * stride;
            dv[u] = 1.0;
            if (i < D.m) {
                if (failed) { if (A_.z_in) dv[u] = __ldg(&A_.z_in[(size_t)i * A_.z_stride]); }
                else dv[u] = __ldcg(&dfin[i]);
            }
        }
#pragma unroll
        for (int u = 0; u < kEpi; ++u) {
            const int i = i0 + u * stride;
            if (i < D.m) {
                if (failed && A_.z_in) dv[u] = 1.0 / dv[u];
                const double o = A_.invert_out ? 1.0 / dv[u] : dv[u];
                A_.out[i] = o;
                zs[0] += o; zm[0] = fmax(zm[0], o); zm[1] = fmax(zm[1], -o);
            }
        }
    }
    if (A_.zstats) block_reduce_store<1, 2>(zs, zm, A_.zstats);
    if (blockIdx.x == 0 && tid == 0) sh->t_abs[1] = globaltimer();
}

}  // namespace rsdsfm
