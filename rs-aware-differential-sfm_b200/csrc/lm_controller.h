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
// The struct is plain data + __host__ __device__ methods so the same code drives the
// host-stepped solver and the persistent on-device solver.
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

// Layout of the reduced sums the per-pixel passes deliver.
// Pass A (evaluation at x + Schur elimination):   sums, then maxima
struct SumsA {
    enum { COST = 0, SUMSQ_D = 1, GF = 2, CSF = GF + kMaxNF, RHS = CSF + kMaxNF, S = RHS + kMaxNF,
           NS = S + kTri };
    enum { GMAX_E = 0, BAD = 1, NM = 2 };
};
// Pass B (back substitution + candidate evaluation)
struct SumsB {
    enum { MCC = 0, STEP_SQ = 1, CAND_COST = 2, NS = 3 };
    enum { BAD_STEP = 0, BAD_CAND = 1, NM = 2 };
};

enum LmNext { LM_RUN_A = 0, LM_RUN_B = 1, LM_DONE = 2 };

struct LmController {
    rsdsfm_lm_options opt;
    int nf;
    // state
    double f[kMaxNF], f_cand[kMaxNF], scale_f[kMaxNF], delta_f[kMaxNF], diag_f[kMaxNF];
    double radius, decrease_factor;
    int iteration, step_is_successful, num_consecutive_invalid, reuse_diagonal;
    int phase;            // 0: first evaluation pending, 1: evaluation after an accepted step pending,
                          // 2: re-elimination at unchanged x (new radius) pending
    double x_cost, gmax, x_norm, cand_cost, rho;
    int accepted_last;    // set by after_B: the candidate became x (caller swaps depth buffers)
    // results
    int termination, reason, num_successful, num_unsuccessful;
    double initial_cost;

    RS_HD void init(const rsdsfm_lm_options &o, int nf_, const double *f0)
    {
        opt = o; nf = nf_;
        for (int j = 0; j < kMaxNF; ++j) { f[j] = j < nf ? f0[j] : 0.0; f_cand[j] = f[j]; scale_f[j] = 1.0; delta_f[j] = 0.0; diag_f[j] = 0.0; }
        radius = o.initial_trust_region_radius; decrease_factor = 2.0;
        iteration = 0; step_is_successful = 0; num_consecutive_invalid = 0; reuse_diagonal = 0;
        phase = 0; x_cost = 0.0; gmax = 0.0; x_norm = 0.0; cand_cost = 0.0; rho = 0.0; accepted_last = 0;
        termination = RSDSFM_NO_CONVERGENCE; reason = RSDSFM_REASON_NONE; num_successful = 0; num_unsuccessful = 0;
        initial_cost = 0.0;
    }

    RS_HD static double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

    RS_HD LmNext finish(int term, int why) { termination = term; reason = why; return LM_DONE; }

    // Dense Cholesky (LL^T) of the nf x nf reduced system, Eigen::LLT semantics: fails when a
    // pivot is not positive (or NaN).  a: full row-major nf x nf.
    RS_HD static bool cholesky_solve(const double *a, int n, const double *b, double *x)
    {
        double l[kMaxNF * kMaxNF];
        for (int i = 0; i < n * n; ++i) l[i] = 0.0;
        for (int j = 0; j < n; ++j) {
            double d = a[j * n + j];
            for (int k = 0; k < j; ++k) d -= l[j * n + k] * l[j * n + k];
            if (!(d > 0.0)) return false;
            d = sqrt(d);
            l[j * n + j] = d;
            for (int i = j + 1; i < n; ++i) {
                double s = a[i * n + j];
                for (int k = 0; k < j; ++k) s -= l[i * n + k] * l[j * n + k];
                l[i * n + j] = s / d;
            }
        }
        double y[kMaxNF];
        for (int i = 0; i < n; ++i) {
            double s = b[i];
            for (int k = 0; k < i; ++k) s -= l[i * n + k] * y[k];
            y[i] = s / l[i * n + i];
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = y[i];
            for (int k = i + 1; k < n; ++k) s -= l[k * n + i] * x[k];
            x[i] = s / l[i * n + i];
        }
        return true;
    }

    // Consumes the sums of a pass A run at the current x with the current radius.
    // Returns LM_RUN_B (delta_f holds the motion step to back-substitute), LM_RUN_A (the reduced
    // solve failed: radius was shrunk, eliminate again) or LM_DONE.
    RS_HD LmNext after_A(const double *sa, const double *ma)
    {
        const bool bad = ma[SumsA::BAD] > 0.0;
        if (phase == 0) {
            // IterationZero: EvaluateGradientAndJacobian + Jacobi scaling fixed here
            if (bad) { initial_cost = 0.0; return finish(RSDSFM_FAILURE, RSDSFM_REASON_EVAL_FAILED); }
            for (int j = 0; j < nf; ++j) scale_f[j] = 1.0 / (1.0 + sqrt(sa[SumsA::CSF + j]));
            x_cost = sa[SumsA::COST];
            initial_cost = x_cost;
        } else if (phase == 1) {
            // HandleSuccessfulStep: evaluation at the new x
            if (bad) return finish(RSDSFM_FAILURE, RSDSFM_REASON_EVAL_FAILED);
            x_cost = sa[SumsA::COST];
            step_is_successful = 1;
        }
        if (phase == 0 || phase == 1) {
            double g = ma[SumsA::GMAX_E], xs = sa[SumsA::SUMSQ_D];
            for (int j = 0; j < nf; ++j) {
                // Ceres: |x - Plus(x, -g)|_inf
                double proj = f[j] + (-sa[SumsA::GF + j]);
                g = fmax(g, fabs(f[j] - proj));
                xs += f[j] * f[j];
            }
            gmax = g;
            x_norm = sqrt(xs);
            // FinalizeIterationAndCheckIfMinimizerCanContinue
            if (iteration > 0) { if (step_is_successful) num_successful++; else num_unsuccessful++; }
        } else {
            if (iteration > 0) num_unsuccessful++;
        }
        if (iteration >= opt.max_num_iterations) return finish(RSDSFM_NO_CONVERGENCE, RSDSFM_REASON_MAX_ITER);
        if (step_is_successful && gmax <= opt.gradient_tolerance) return finish(RSDSFM_CONVERGENCE, RSDSFM_REASON_GRADIENT_TOL);
        if (radius <= opt.min_trust_region_radius) return finish(RSDSFM_CONVERGENCE, RSDSFM_REASON_MIN_RADIUS);
        iteration++;
        step_is_successful = 0;
        accepted_last = 0;

        // LevenbergMarquardtStrategy::ComputeStep on the Schur-reduced system
        bool ok = true;
        if (nf > 0) {
            if (!reuse_diagonal)
                for (int j = 0; j < nf; ++j)
                    diag_f[j] = clampd(sa[SumsA::CSF + j] * scale_f[j] * scale_f[j], opt.min_lm_diagonal, opt.max_lm_diagonal);
            double lhs[kMaxNF * kMaxNF], rhs[kMaxNF], y[kMaxNF];
            for (int i = 0; i < nf; ++i) {
                rhs[i] = sa[SumsA::RHS + i] * scale_f[i];
                for (int j = i; j < nf; ++j) {
                    double vv = sa[SumsA::S + tri_index(nf, i, j)] * scale_f[i] * scale_f[j];
                    lhs[i * nf + j] = vv; lhs[j * nf + i] = vv;
                }
                double Df = sqrt(diag_f[i] / radius);
                lhs[i * nf + i] += Df * Df;
            }
            ok = cholesky_solve(lhs, nf, rhs, y);
            for (int j = 0; j < nf && ok; ++j) {
                if (!isfinite(y[j])) ok = false;
                delta_f[j] = -y[j] * scale_f[j];     // step = -y ; delta = step o scale
            }
        }
        reuse_diagonal = 1;
        if (!ok) return invalid_step();
        return LM_RUN_B;
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
        return LM_RUN_A;
    }

    // Consumes the sums of a pass B (candidate point).  Returns LM_RUN_A or LM_DONE.
    RS_HD LmNext after_B(const double *sb, const double *mb)
    {
        const double model_cost_change = -sb[SumsB::MCC];
        const bool step_finite = !(mb[SumsB::BAD_STEP] > 0.0);
        if (!(step_finite && model_cost_change > 0.0)) return invalid_step();
        num_consecutive_invalid = 0;
        double step_sq = sb[SumsB::STEP_SQ];
        for (int j = 0; j < nf; ++j) {
            f_cand[j] = f[j] + delta_f[j];
            double dd = f[j] - f_cand[j];
            step_sq += dd * dd;
        }
        cand_cost = (mb[SumsB::BAD_CAND] > 0.0) ? DBL_MAX : sb[SumsB::CAND_COST];
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
        } else {
            radius = radius / decrease_factor;
            decrease_factor *= 2.0;
            reuse_diagonal = 1;
            phase = 2;
        }
        return LM_RUN_A;
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
