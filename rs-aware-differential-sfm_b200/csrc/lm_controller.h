// lm_controller.h -- the O(1) part of the Levenberg-Marquardt refinement: trust-region control,
// the tiny dense Cholesky solve of the Schur-reduced motion system and the termination tests.
//
// It reproduces the behaviour of Ceres 1.14 as the reference configures it
// (nonlinearRefinement.cc:159-163, :224-228: Solver::Options defaults + DENSE_SCHUR):
// TrustRegionMinimizer::Minimize, LevenbergMarquardtStrategy::{ComputeStep,StepAccepted,
// StepRejected}, TrustRegionStepEvaluator (monotonic steps) and DenseSchurComplementSolver's
// reduced solve.  Ceres is a third-party dependency of the reference (README.md:32-33, version
// 1.14.0) that is not vendored; the rules below restate its published algorithm (SURVEY.md
// Appendix B).  The O(n) work (residuals, Jacobians, Schur elimination, back substitution) is
// done by the kernels in refine.cu / ransac.cu, which hand this controller their reduced sums.
//
// Radius-independent Schur sums.  With e = dr/dd (2-vector), n = (-e1, e0) and F = dr/d(motion),
// the per-pixel projector of the 1x1 e-block elimination is
//     I - q e e^T ,  q = s_e^2 / (s_e^2 e^Te + clamp(s_e^2 e^Te, min, max) / radius)
// and for every pixel whose LM diagonal is not clamped it equals
//     (n n^T + e e^T / (radius + 1)) / e^Te .
// The passes therefore accumulate   G1 = sum (F^Tn)(n^TF)/e^Te,  G2 = sum (F^Te)(e^TF)/e^Te,
// h1 = sum (F^Tn)(n^Tr)/e^Te,  h2 = sum (F^Te)(e^Tr)/e^Te   once per evaluation point, and
//     S = G1 + G2/(radius+1),  rhs = h1 + h2/(radius+1),  F^TF = G1 + G2,  F^Tr = h1 + h2
// follow for ANY radius: a rejected step needs no new pass over the pixels.  The few pixels
// whose diagonal IS clamped (|e| ~ 0: at the focus of expansion) put F^TF / F^Tr into G1 / h1
// and are listed as exceptions; their radius-dependent term is subtracted by the controller.
//
// Plain data + __host__ __device__ methods: the same code drives the persistent on-device
// solver (refine.cu) and the host-stepped RANSAC depth solver (ransac.cu).
#pragma once

#include <float.h>
#include <math.h>

#include "../../include/rsdsfm.h"

#ifdef __CUDACC__
#define RS_HD __host__ __device__
#else
#define RS_HD
#endif

namespace rsdsfm {

constexpr int kMaxNF = 7;                          // v(3) + w(3) + k(1)
constexpr int kTri = kMaxNF * (kMaxNF + 1) / 2;    // packed upper triangle

RS_HD inline int tri_index(int nf, int i, int j)   // i <= j
{
    return i * nf - (i * (i - 1)) / 2 + (j - i);
}

// Reduced sums of an evaluation pass (pass A) at the current point; radius independent.
struct EvalSums {
    double cost;       // sum 0.5 |r|^2
    double sumsq_d;    // sum d^2
    double gmax_e;     // max |e^T r|   (gradient of the depth blocks)
    double bad;        // > 0: a residual / Jacobian entry was not finite
    double ee_max;     // max e^Te (first evaluation: bounds the depth-column Jacobi scales from below)
    double G1[kTri], G2[kTri], h1[kMaxNF], h2[kMaxNF];
};
// Radius-dependent correction of the clamped pixels: sum q fe fe^T and sum q fe (e^T r).
struct ExcSums {
    double S[kTri], rhs[kMaxNF];
};
// Reduced sums of a candidate pass (pass B).
struct CandSums {
    double mcc;        // sum (J delta)^T (r + J delta / 2)      (= -model_cost_change)
    double step_sq;    // sum (d - d_cand)^2
    double cand_cost;  // sum 0.5 |r(x_cand)|^2
    double bad_step;   // > 0: a depth step was not finite
    double bad_cand;   // > 0: a candidate residual was not finite
};

enum LmNext { LM_RUN_A = 0, LM_RUN_B = 1, LM_DONE = 2, LM_SOLVE = 3 };

struct LmController {
    rsdsfm_lm_options opt;
    int nf;
    // state
    double f[kMaxNF], f_cand[kMaxNF], scale_f[kMaxNF], delta_f[kMaxNF], diag_f[kMaxNF];
    double radius, decrease_factor;
    int iteration, step_is_successful, num_consecutive_invalid, reuse_diagonal;
    int phase;            // 0: first evaluation pending, 1: evaluation after an accepted step pending,
                          // 2: same point, new radius
    double x_cost, gmax, x_norm, cand_cost, rho;
    int accepted_last;    // set by on_candidate: the candidate became x (depth buffers swap)
    EvalSums ev;          // sums of the last evaluation pass (kept for re-solves at a new radius)
    double ee_fast_min;   // e^Te >= this  =>  the pixel's LM diagonal is certainly not clamped from below
    // results
    int termination, reason, num_successful, num_unsuccessful;
    double initial_cost;

    RS_HD void init(const rsdsfm_lm_options &o, int nf_, const double *f0)
    {
        opt = o; nf = nf_;
        for (int j = 0; j < kMaxNF; ++j) { f[j] = (j < nf && f0) ? f0[j] : 0.0; f_cand[j] = f[j]; scale_f[j] = 1.0; delta_f[j] = 0.0; diag_f[j] = 0.0; }
        radius = o.initial_trust_region_radius; decrease_factor = 2.0;
        iteration = 0; step_is_successful = 0; num_consecutive_invalid = 0; reuse_diagonal = 0;
        phase = 0; x_cost = 0.0; gmax = 0.0; x_norm = 0.0; cand_cost = 0.0; rho = 0.0; accepted_last = 0;
        termination = RSDSFM_NO_CONVERGENCE; reason = RSDSFM_REASON_NONE; num_successful = 0; num_unsuccessful = 0;
        initial_cost = 0.0;
        ev.cost = ev.sumsq_d = ev.gmax_e = ev.bad = ev.ee_max = 0.0;
        ee_fast_min = 0.0;
        for (int j = 0; j < kTri; ++j) { ev.G1[j] = 0.0; ev.G2[j] = 0.0; }
        for (int j = 0; j < kMaxNF; ++j) { ev.h1[j] = 0.0; ev.h2[j] = 0.0; }
    }

    RS_HD static double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

    RS_HD LmNext finish(int term, int why) { termination = term; reason = why; return LM_DONE; }

    // Consumes the sums of an evaluation pass run at the current x (phase 0 or 1).
    // Returns LM_SOLVE (call solve_step with the exception sums at `radius`) or LM_DONE.
    RS_HD LmNext on_eval(const EvalSums &e_in)
    {
        ev = e_in;
        return on_eval_stored();
    }

    // Same, with the sums already written into this->ev (the device controller fills it in place).
    RS_HD LmNext on_eval_stored()
    {
        const EvalSums &e = ev;
        const bool bad = e.bad > 0.0;
        if (phase == 0) {
            // IterationZero: EvaluateGradientAndJacobian; the Jacobi scaling is fixed here
            if (bad) { initial_cost = 0.0; return finish(RSDSFM_FAILURE, RSDSFM_REASON_EVAL_FAILED); }
            for (int j = 0; j < nf; ++j) {
                const int t = tri_index(nf, j, j);
                scale_f[j] = 1.0 / (1.0 + sqrt(e.G1[t] + e.G2[t]));
            }
            x_cost = e.cost;
            initial_cost = x_cost;
            // s_e = 1/(1+|e(x0)|) >= 1/(1+sqrt(ee_max)):  e^Te >= min_diag (1+sqrt(ee_max))^2  =>  s_e^2 e^Te >= min_diag
            const double t = 1.0 + sqrt(e.ee_max);
            ee_fast_min = opt.min_lm_diagonal * t * t;
            // IterationZero ends with step_is_valid = step_is_successful = true: iteration 0 counts as a
            // successful step and the gradient tolerance is tested before the first step is computed
            step_is_successful = 1;
        } else {
            // HandleSuccessfulStep: evaluation at the new x
            if (bad) return finish(RSDSFM_FAILURE, RSDSFM_REASON_EVAL_FAILED);
            x_cost = e.cost;
            step_is_successful = 1;
        }
        double g = e.gmax_e, xs = e.sumsq_d;
        for (int j = 0; j < nf; ++j) {
            // Ceres: |x - Plus(x, -g)|_inf with g = F^T r = h1 + h2
            const double proj = f[j] + (-(e.h1[j] + e.h2[j]));
            g = fmax(g, fabs(f[j] - proj));
            xs += f[j] * f[j];
        }
        gmax = g;
        x_norm = sqrt(xs);
        return begin_iteration();
    }

    // FinalizeIterationAndCheckIfMinimizerCanContinue + start of the next iteration.
    RS_HD LmNext begin_iteration()
    {
        if (step_is_successful) num_successful++; else num_unsuccessful++;
        if (iteration >= opt.max_num_iterations) return finish(RSDSFM_NO_CONVERGENCE, RSDSFM_REASON_MAX_ITER);
        if (step_is_successful && gmax <= opt.gradient_tolerance) return finish(RSDSFM_CONVERGENCE, RSDSFM_REASON_GRADIENT_TOL);
        if (radius <= opt.min_trust_region_radius) return finish(RSDSFM_CONVERGENCE, RSDSFM_REASON_MIN_RADIUS);
        iteration++;
        step_is_successful = 0;
        accepted_last = 0;
        return LM_SOLVE;
    }

    // LevenbergMarquardtStrategy::ComputeStep on the Schur-reduced system at the current radius.
    // exc: clamped-pixel correction evaluated at this radius (may be NULL when there is none).
    // Returns LM_RUN_B (delta_f = motion step to back-substitute), LM_SOLVE (the solve failed, the
    // radius was shrunk: call again with the exception sums at the new radius) or LM_DONE.
    RS_HD LmNext solve_step(const ExcSums *exc)
    {
        bool ok = true;
        if (nf == 6) ok = solve_fixed<6>(exc);
        else if (nf == 7) ok = solve_fixed<7>(exc);
        reuse_diagonal = 1;
        if (!ok) return invalid_step();
        return LM_RUN_B;
    }

    // Compose S(radius), add the LM diagonal, Cholesky-factorise (Eigen::LLT semantics: fails on a
    // non-positive or NaN pivot) and solve.  N is a compile-time constant so that every loop
    // unrolls and the N x N system lives in registers (this runs on ONE thread of the last CTA,
    // with the whole grid waiting for it).
    template <int N>
    RS_HD bool solve_fixed(const ExcSums *exc)
    {
        double sc[N], l[N][N], y[N];
#pragma unroll
        for (int j = 0; j < N; ++j) sc[j] = scale_f[j];
        if (!reuse_diagonal) {
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const int t = j * N - (j * (j - 1)) / 2;
                diag_f[j] = clampd((ev.G1[t] + ev.G2[t]) * sc[j] * sc[j], opt.min_lm_diagonal, opt.max_lm_diagonal);
            }
        }
        const double eps = 1.0 / (radius + 1.0);
        // lower triangle of the scaled, damped system (l[i][j], i >= j), then in-place LL^T
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double rv = ev.h1[i] + ev.h2[i] * eps;
            if (exc) rv -= exc->rhs[i];
            y[i] = rv * sc[i];
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                const int t = j * N - (j * (j - 1)) / 2 + (i - j);      // tri_index(N, j, i)
                double sv = ev.G1[t] + ev.G2[t] * eps;
                if (exc) sv -= exc->S[t];
                l[i][j] = sv * (sc[i] * sc[j]);
            }
            const double Df = sqrt(diag_f[i] / radius);
            l[i][i] += Df * Df;
        }
        bool ok = true;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double d = l[j][j];
#pragma unroll
            for (int k = 0; k < j; ++k) d -= l[j][k] * l[j][k];
            if (!(d > 0.0)) ok = false;
            d = sqrt(d);
            l[j][j] = d;
            const double inv = 1.0 / d;
#pragma unroll
            for (int i = j + 1; i < N; ++i) {
                double s = l[i][j];
#pragma unroll
                for (int k = 0; k < j; ++k) s -= l[i][k] * l[j][k];
                l[i][j] = s * inv;
            }
        }
        if (!ok) return false;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double s = y[i];
#pragma unroll
            for (int k = 0; k < i; ++k) s -= l[i][k] * y[k];
            y[i] = s / l[i][i];
        }
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {
            double s = y[i];
#pragma unroll
            for (int k = i + 1; k < N; ++k) s -= l[k][i] * y[k];
            y[i] = s / l[i][i];
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            if (!isfinite(y[j])) ok = false;
            delta_f[j] = -y[j] * sc[j];          // step = -y ; delta = step o scale
        }
        return ok;
    }

    RS_HD LmNext invalid_step()
    {
        // HandleInvalidStep
        num_consecutive_invalid++;
        if (num_consecutive_invalid >= opt.max_num_consecutive_invalid_steps) {
            num_unsuccessful++;
            return finish(RSDSFM_FAILURE, RSDSFM_REASON_INVALID_STEPS);
        }
        radius = radius / decrease_factor;   // StepIsInvalid -> StepRejected(0)
        decrease_factor *= 2.0;
        reuse_diagonal = 1;
        phase = 2;
        return begin_iteration();
    }

    // Consumes the sums of a candidate pass.  Returns LM_RUN_A (accepted: evaluate at the new x),
    // LM_SOLVE (rejected / invalid: same x, smaller radius) or LM_DONE.
    RS_HD LmNext on_candidate(const CandSums &c)
    {
        const double model_cost_change = -c.mcc;
        const bool step_finite = !(c.bad_step > 0.0);
        if (!(step_finite && model_cost_change > 0.0)) return invalid_step();
        num_consecutive_invalid = 0;
        double step_sq = c.step_sq;
        for (int j = 0; j < nf; ++j) {
            f_cand[j] = f[j] + delta_f[j];
            const double dd = f[j] - f_cand[j];
            step_sq += dd * dd;
        }
        cand_cost = (c.bad_cand > 0.0) ? DBL_MAX : c.cand_cost;
        // ParameterToleranceReached (tested on the candidate, which is then discarded)
        const double step_norm = sqrt(step_sq);
        if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance))
            return finish(RSDSFM_CONVERGENCE, RSDSFM_REASON_PARAMETER_TOL);
        // FunctionToleranceReached
        const double cost_change = x_cost - cand_cost;
        if (fabs(cost_change) <= opt.function_tolerance * x_cost)
            return finish(RSDSFM_CONVERGENCE, RSDSFM_REASON_FUNCTION_TOL);
        // IsStepSuccessful
        rho = cost_change / model_cost_change;
        if (rho > opt.min_relative_decrease) {
            for (int j = 0; j < nf; ++j) f[j] = f_cand[j];
            accepted_last = 1;
            const double t = 2.0 * rho - 1.0;
            radius = radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
            radius = fmin(opt.max_trust_region_radius, radius);
            decrease_factor = 2.0;
            reuse_diagonal = 0;
            phase = 1;
            return LM_RUN_A;
        }
        radius = radius / decrease_factor;
        decrease_factor *= 2.0;
        reuse_diagonal = 1;
        phase = 2;
        return begin_iteration();
    }

    RS_HD void fill_summary(rsdsfm_lm_summary *s) const
    {
        s->termination = termination; s->reason = reason; s->iterations = iteration;
        s->num_successful = num_successful; s->num_unsuccessful = num_unsuccessful;
        s->initial_cost = initial_cost;
        s->final_cost = (termination == RSDSFM_FAILURE) ? initial_cost : x_cost;
        s->final_radius = radius; s->final_gradient_max_norm = gmax; s->device_ms = 0.0;
    }
};

}  // namespace rsdsfm
