/*
 * rsdsfm_oracle.c -- CPU restatement of the dense optimisation core of
 * ThomasZiegler/RS-aware-differential-SfM (reference mounted at /root/reference, paths below
 * are relative to its src/ directory).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker / the reported CPU
 * baseline.  The product (rs-aware-differential-sfm_b200/) never links, imports or executes it.
 *
 * PARITY UNPINNED.  The reference ships no tests, golden vectors or fixtures for this path, its
 * example data tarballs are absent (.MISSING_LARGE_BLOBS) and it cannot be compiled here: every
 * source includes Eigen 3.3.4 / Ceres 1.14.0 / OpenCV 3.4.0 headers, none of which is installed
 * (and none of which is vendored in the reference tree).  The arithmetic that lives in those
 * third-party libraries is restated from their published algorithms:
 *   - Ceres 1.14 TrustRegionMinimizer + LevenbergMarquardtStrategy + DENSE_SCHUR
 *     (SchurEliminator / DenseSchurComplementSolver), defaults of Solver::Options 1.14
 *     => orc_lm_solve() below.
 *   - Eigen decompositions => orc_linalg.h.
 * What substitutes for golden vectors is listed in DESIGN.md (analytic known-answer tests,
 * LAPACK cross-checks of the decompositions, Schur == dense normal equations, an independent
 * numpy restatement of the LM loop).
 *
 * Conventions: Eigen Array2Xd / Matrix2Xd = interleaved pairs [x0,y0,x1,y1,...]; Array3Xd =
 * interleaved triples; Eigen MatrixXd(rows,cols) depth map = column-major (y + x*rows);
 * cv::Mat 8UC3 = row-major interleaved BGR; cv::Mat_<Point2d> flow = row-major (dx,dy).
 * Build with -ffp-contract=off: the CUDA RANSAC scoring path mirrors the per-point operation
 * sequence below with explicit round-to-nearest intrinsics, so that inlier sets are bit-exact.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "orc_linalg.h"


#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* a2: flatten + normalise glue                                    main.cc:398-432,            */
/*                                                                 errorMeasure.cpp:66-97      */
/* ------------------------------------------------------------------------------------------ */
/* flow_img: rows x cols x 2 row-major doubles (cv::Mat_<Point_<double>>).
 * Output arrays hold rows*cols points and are pre-filled the way the reference constructs them
 * (coord = ones, flow = zeros); the first `n` (return value) entries are the kept pixels in
 * COLUMN-MAJOR pixel order.  errorMeasure.cpp truncates to n (conservativeResize :96-97),
 * main.cc keeps the padded tail (Q3) -- the caller chooses by how many points it passes on. */
ORC_API int orc_flatten(const double *flow_img, int rows, int cols, double fx, double fy, double cx,
                        double cy, double gamma, double flow_threshold, double *coord, double *flow,
                        double *coord_px, double *flow_px)
{
    long total = (long)rows * cols;
    for (long p = 0; p < total; ++p) {
        coord[2 * p] = 1.0; coord[2 * p + 1] = 1.0;
        coord_px[2 * p] = 1.0; coord_px[2 * p + 1] = 1.0;
        flow[2 * p] = 0.0; flow[2 * p + 1] = 0.0;
        flow_px[2 * p] = 0.0; flow_px[2 * p + 1] = 0.0;
    }
    int position = 0;
    for (int i = 0; i < cols; ++i) {
        for (int j = 0; j < rows; ++j) {
            double dx = flow_img[2 * ((long)j * cols + i)];
            double dy = flow_img[2 * ((long)j * cols + i) + 1];
            double norm = dx * dx + dy * dy;
            if (norm > flow_threshold) {
                coord_px[2 * position] = (double)i;
                coord_px[2 * position + 1] = (double)j;
                flow_px[2 * position] = dx;
                flow_px[2 * position + 1] = dy;
                flow[2 * position] = dx * gamma / fx;
                flow[2 * position + 1] = dy * gamma / fy;
                coord[2 * position] = (i - cx) * 1.0 / fx;
                coord[2 * position + 1] = (j - cy) * 1.0 / fy;
                position++;
            }
        }
    }
    return position;
}

/* a3: minimal::getAlpha                                           minimal.cc:179-186 */
ORC_API void orc_get_alpha(const double *flow_px, int n, double h, double gamma, double *alpha)
{
    for (int i = 0; i < n; ++i) alpha[i] = 1 + gamma * flow_px[2 * i + 1] / h;
}

/* a4: minimal::getAlphaK                                          minimal.cc:188-197 */
ORC_API void orc_get_alpha_k(const double *q_px, const double *flow_px, int n, double h, double gamma,
                             double *alpha_k)
{
    for (int i = 0; i < n; ++i) {
        double part1 = gamma * q_px[2 * i + 1] / h;
        double part2 = 1.0 + gamma * (q_px[2 * i + 1] + flow_px[2 * i + 1]) / h;
        alpha_k[i] = 0.5 * (part2 * part2 - part1 * part1);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a5: minimal::calculateVelocities                                minimal.cc:36-177           */
/* ------------------------------------------------------------------------------------------ */
static void mat3_mul(const double *a, const double *b, double *c)
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            c[i * 3 + j] = a[i * 3 + 0] * b[0 * 3 + j] + a[i * 3 + 1] * b[1 * 3 + j] + a[i * 3 + 2] * b[2 * 3 + j];
}
static void mat3_t(const double *a, double *t)
{
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) t[i * 3 + j] = a[j * 3 + i];
}
static void rot_y(double ang, double *r)
{   /* Eigen::AngleAxisd(ang, UnitY()).toRotationMatrix() */
    double c = cos(ang), s = sin(ang);
    r[0] = c; r[1] = 0; r[2] = s; r[3] = 0; r[4] = 1; r[5] = 0; r[6] = -s; r[7] = 0; r[8] = c;
}
static void rot_z(double ang, double *r)
{
    double c = cos(ang), s = sin(ang);
    r[0] = c; r[1] = -s; r[2] = 0; r[3] = s; r[4] = c; r[5] = 0; r[6] = 0; r[7] = 0; r[8] = 1;
}
/* m * r * sig * m^T */
static void sandwich(const double *m, const double *r, const double *sig, double *out)
{
    double t1[9], t2[9], mt[9];
    mat3_mul(m, r, t1);
    mat3_mul(t1, sig, t2);
    mat3_t(m, mt);
    mat3_mul(t2, mt, out);
}

/* q,u: 2x9 interleaved; alpha, alpha_k: 9; out7 = (w[3], v[3], k).
 * svd_sign: +1 keeps the null vector as computed, -1 negates it (the sign Eigen's JacobiSVD
 * would return is unknowable here; the pipeline is invariant to it after the mean-depth sign
 * fix, Q8 -- tests exercise both).  evec_flip: bitmask, bit j negates eigenvector column j of
 * the 3x3 symmetric decomposition (same reason).  Pass (1, 0) for the default. */
ORC_API void orc_calculate_velocities_ex(const double *q, const double *u, const double *alpha,
                                         const double *alpha_k, int use_alpha_k, int svd_sign,
                                         int evec_flip, double *out7)
{
    const double THRESHOLD_LAMBDA = 0.000001;
    const double TOL_IMAG = 0.00001;
    const int n = 9;
    double k = 0;
    double beta[9];
    double z[81]; /* n x 9 row-major */
    for (int i = 0; i < n; ++i) {
        double qx = q[2 * i], qy = q[2 * i + 1], ux = u[2 * i], uy = u[2 * i + 1];
        beta[i] = 1.0;
        z[i * 9 + 0] = -uy;
        z[i * 9 + 1] = ux;
        z[i * 9 + 2] = uy * qx - ux * qy;
        z[i * 9 + 3] = qx * qx;
        z[i * 9 + 4] = 2.0 * qx * qy;
        z[i * 9 + 5] = 2.0 * qx;
        z[i * 9 + 6] = qy * qy;
        z[i * 9 + 7] = 2 * qy;
        z[i * 9 + 8] = 1.0;
    }
    if (use_alpha_k) {
        double a[9], a_inv[9], efhj[36], dg[18], bc[18];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i * 3 + j] = z[i * 9 + j];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) efhj[i * 6 + j] = z[(3 + i) * 9 + 3 + j];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 3; ++j) dg[i * 3 + j] = z[(3 + i) * 9 + j];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 6; ++j) bc[i * 6 + j] = z[i * 9 + 3 + j];
        orc_inverse(a, 3, a_inv);
        /* dga = dg * a_inv (6x3) */
        double dga[18];
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 3; ++j)
                dga[i * 3 + j] = dg[i * 3 + 0] * a_inv[0 * 3 + j] + dg[i * 3 + 1] * a_inv[1 * 3 + j] +
                                 dg[i * 3 + 2] * a_inv[2 * 3 + j];
        double p[36], pk[36];
        for (int i = 0; i < 6; ++i) {
            for (int j = 0; j < 6; ++j) {
                double s = 0.0, sk = 0.0;
                for (int c = 0; c < 3; ++c) {
                    s += dga[i * 3 + c] * alpha[c] * bc[c * 6 + j];
                    sk += dga[i * 3 + c] * alpha_k[c] * bc[c * 6 + j];
                }
                p[i * 6 + j] = alpha[3 + i] * efhj[i * 6 + j] - s;
                pk[i * 6 + j] = alpha_k[3 + i] * efhj[i * 6 + j] - sk;
            }
        }
        double pk_inv[36], m[36];
        orc_inverse(pk, 6, pk_inv);
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) {
                double s = 0.0;
                for (int c = 0; c < 6; ++c) s += p[i * 6 + c] * pk_inv[c * 6 + j];
                m[i * 6 + j] = s;
            }
        double wr[6], wi[6];
        orc_eigvals_general(m, 6, wr, wi);
        k = INFINITY;
        for (int i = 0; i < 6; ++i) {
            if ((fabs(wi[i]) < TOL_IMAG) && fabs(wr[i]) < fabs(k)) k = wr[i];
        }
        for (int i = 0; i < n; ++i) beta[i] = (alpha[i] + k * alpha_k[i]) * (2.0 / (2.0 + k));
    } else {
        for (int i = 0; i < n; ++i) beta[i] = alpha[i];
    }
    for (int i = 0; i < n; ++i)
        for (int c = 3; c < 9; ++c) z[i * 9 + c] *= beta[i];

    /* Step 1: e = right singular vector of the smallest singular value */
    double e[9];
    {
        double zz[81], v[81], sig[9];
        int finite = 1;
        memcpy(zz, z, sizeof zz);
        for (int i = 0; i < 81; ++i) if (!isfinite(zz[i])) finite = 0;
        if (finite) {
            orc_jacobi_svd(zz, 9, 9, v, sig);
            int best = 0;
            for (int j = 1; j < 9; ++j) if (sig[j] < sig[best]) best = j;
            for (int i = 0; i < 9; ++i) e[i] = v[i * 9 + best] * (double)svd_sign;
        } else {
            for (int i = 0; i < 9; ++i) e[i] = NAN;
        }
    }
    double norm_v0 = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    for (int i = 0; i < 9; ++i) e[i] = e[i] / norm_v0;
    double v0[3] = {e[0], e[1], e[2]};
    double s[9] = {e[3], e[4], e[5], e[4], e[6], e[7], e[5], e[7], e[8]};

    /* Step 2 */
    double lamb[3], v1[9];
    {
        double sc[9];
        int finite = 1;
        memcpy(sc, s, sizeof sc);
        for (int i = 0; i < 9; ++i) if (!isfinite(sc[i])) finite = 0;
        if (finite) {
            orc_sym_eig(sc, 3, lamb, v1);
        } else {
            for (int i = 0; i < 3; ++i) lamb[i] = NAN;
            for (int i = 0; i < 9; ++i) v1[i] = NAN;
        }
        for (int j = 0; j < 3; ++j)
            if (evec_flip & (1 << j))
                for (int r = 0; r < 3; ++r) v1[r * 3 + j] = -v1[r * 3 + j];
    }
    for (int r = 0; r < 3; ++r) { double t = v1[r * 3 + 0]; v1[r * 3 + 0] = v1[r * 3 + 2]; v1[r * 3 + 2] = t; }
    double sigma[3];
    sigma[0] = (2 * lamb[2] + lamb[1] - lamb[0]) / 3;
    sigma[1] = (lamb[2] + 2 * lamb[1] + lamb[0]) / 3;
    sigma[2] = (-lamb[2] + lamb[1] + 2 * lamb[0]) / 3;
    /* Step 3 */
    double lambda = sigma[0] - sigma[2];
    double theta = 0;
    if (lambda < THRESHOLD_LAMBDA) {
        /* reference only prints a warning here */
    } else {
        theta = acos(-sigma[1] / lambda);
    }
    double r_v[9], r_u[9], r_vt[9], v_[9], u_[9];
    rot_y((theta - M_PI) / 2, r_v);
    rot_y(theta, r_u);
    mat3_t(r_v, r_vt);
    mat3_mul(v1, r_vt, v_);
    {
        double nv[9];
        for (int i = 0; i < 9; ++i) nv[i] = -v_[i];
        mat3_mul(nv, r_u, u_);
    }
    double sig1[9] = {1, 0, 0, 0, 1, 0, 0, 0, 0};
    double sig_lamb[9];
    for (int i = 0; i < 9; ++i) sig_lamb[i] = lambda * sig1[i];
    double r_z1[9], r_z2[9];
    rot_z(M_PI / 2, r_z1);
    rot_z(-M_PI / 2, r_z2);
    double vh[4][9];
    sandwich(v_, r_z1, sig1, vh[0]);
    sandwich(v_, r_z2, sig1, vh[1]);
    sandwich(u_, r_z1, sig1, vh[2]);
    sandwich(u_, r_z2, sig1, vh[3]);
    /* Step 4 */
    double dotv[4];
    for (int c = 0; c < 4; ++c) {
        double a0 = vh[c][2 * 3 + 1], a1 = vh[c][0 * 3 + 2], a2 = vh[c][1 * 3 + 0];
        dotv[c] = a0 * v0[0] + a1 * v0[1] + a2 * v0[2];
    }
    int index_max = 0;
    for (int c = 1; c < 4; ++c) if (dotv[c] > dotv[index_max]) index_max = c; /* maxCoeff: first max */
    double w_hat[9];
    switch (index_max) {
        case 0: sandwich(u_, r_z1, sig_lamb, w_hat); break;
        case 1: sandwich(u_, r_z2, sig_lamb, w_hat); break;
        case 2: sandwich(v_, r_z1, sig_lamb, w_hat); break;
        default: sandwich(v_, r_z2, sig_lamb, w_hat); break;
    }
    out7[0] = w_hat[2 * 3 + 1];
    out7[1] = w_hat[0 * 3 + 2];
    out7[2] = w_hat[1 * 3 + 0];
    out7[3] = v0[0]; out7[4] = v0[1]; out7[5] = v0[2];
    out7[6] = k;
}

ORC_API void orc_calculate_velocities(const double *q, const double *u, const double *alpha,
                                      const double *alpha_k, int use_alpha_k, double *out7)
{
    orc_calculate_velocities_ex(q, u, alpha, alpha_k, use_alpha_k, 1, 0, out7);
}

/* ------------------------------------------------------------------------------------------ */
/* a7: RsResidual::operator()                                      nonlinearRefinement.cc:32-52 */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double x, y, ux, uy, alpha, alpha_k;
} OrcObs;

static inline void rs_residual(const OrcObs *o, const double *v, const double *w, double k, double d,
                               double *r)
{
    double beta = (2.0 / (2.0 + k)) * (o->alpha + k * o->alpha_k);
    double nb = beta * -1.0;
    double p0 = nb * (d * (o->x * v[2] - v[0]) + (o->x * o->y * w[0]) - (1.0 + o->x * o->x) * w[1] + o->y * w[2]);
    double p1 = nb * (d * (o->y * v[2] - v[1]) + (1.0 + o->y * o->y) * w[0] - o->x * o->y * w[1] - o->x * w[2]);
    r[0] = o->ux - p0;
    r[1] = o->uy - p1;
}

/* Residual + analytic Jacobian (what Ceres' AutoDiffCostFunction<RsResidual,2,3,3,1,1> yields,
 * up to rounding in the motion columns; the depth column is bit-identical).
 * E[2] = dr/dd.  F[2][7] = dr/d(vx,vy,vz,wx,wy,wz,k). */
static inline void rs_residual_jac(const OrcObs *o, const double *v, const double *w, double k, double d,
                                   double *r, double *E, double F[2][7])
{
    double x = o->x, y = o->y;
    double c2 = 2.0 / (2.0 + k);
    double ak = o->alpha + k * o->alpha_k;
    double beta = c2 * ak;
    double nb = beta * -1.0;
    double g0 = x * v[2] - v[0];
    double g1 = y * v[2] - v[1];
    double xy = x * y;
    double ex0 = d * g0 + (xy * w[0]) - (1.0 + x * x) * w[1] + y * w[2];
    double ex1 = d * g1 + (1.0 + y * y) * w[0] - xy * w[1] - x * w[2];
    r[0] = o->ux - nb * ex0;
    r[1] = o->uy - nb * ex1;
    E[0] = beta * g0;
    E[1] = beta * g1;
    double bd = beta * d;
    F[0][0] = -bd;       F[0][1] = 0.0;       F[0][2] = bd * x;
    F[1][0] = 0.0;       F[1][1] = -bd;       F[1][2] = bd * y;
    F[0][3] = beta * xy;             F[0][4] = -(beta * (1.0 + x * x));  F[0][5] = beta * y;
    F[1][3] = beta * (1.0 + y * y);  F[1][4] = -(beta * xy);             F[1][5] = -(beta * x);
    /* d beta / d k = c2 * (alpha_k - (alpha + k alpha_k)/(2+k)) */
    double dbeta = c2 * (o->alpha_k - ak / (2.0 + k));
    F[0][6] = dbeta * ex0;
    F[1][6] = dbeta * ex1;
}

/* ------------------------------------------------------------------------------------------ */
/* Ceres 1.14 trust-region LM with DENSE_SCHUR, restated.               */
/* Free parameter blocks: every inverse depth d_i (the e-blocks, 1x1), and optionally the      */
/* f-blocks v(3), w(3) [free_motion] and k(1) [free_k].  Constant blocks are removed exactly   */
/* like Ceres' preprocessor does (their Jacobian columns are never formed or checked).         */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int max_num_iterations;            /* 50    */
    double function_tolerance;         /* 1e-6  */
    double gradient_tolerance;         /* 1e-10 */
    double parameter_tolerance;        /* 1e-8  */
    double initial_trust_region_radius;/* 1e4   */
    double max_trust_region_radius;    /* 1e16  */
    double min_trust_region_radius;    /* 1e-32 */
    double min_relative_decrease;      /* 1e-3  */
    double min_lm_diagonal;            /* 1e-6  */
    double max_lm_diagonal;            /* 1e32  */
    int max_num_consecutive_invalid_steps; /* 5 */
} OrcLmOptions;

enum { ORC_CONVERGENCE = 0, ORC_NO_CONVERGENCE = 1, ORC_FAILURE = 2 };
enum {
    ORC_REASON_NONE = 0,
    ORC_REASON_PARAMETER_TOL = 1,
    ORC_REASON_FUNCTION_TOL = 2,
    ORC_REASON_GRADIENT_TOL = 3,
    ORC_REASON_MAX_ITER = 4,
    ORC_REASON_MIN_RADIUS = 5,
    ORC_REASON_INVALID_STEPS = 6,
    ORC_REASON_EVAL_FAILED = 7,
    ORC_REASON_NONFINITE_INPUT = 8
};

typedef struct {
    int termination;        /* ORC_CONVERGENCE / NO_CONVERGENCE / FAILURE */
    int reason;
    int iterations;         /* last iteration index (0 = only the initial evaluation) */
    int num_successful;
    int num_unsuccessful;   /* rejected + invalid */
    double initial_cost;
    double final_cost;
    double final_radius;
    double final_gradient_max_norm;
    /* per-iteration trace (first 64 iterations): cost after the iteration, radius, rho */
    double trace_cost[64];
    double trace_radius[64];
    double trace_rho[64];
    int trace_accepted[64];
} OrcLmSummary;

ORC_API void orc_lm_default_options(OrcLmOptions *o)
{
    o->max_num_iterations = 50;
    o->function_tolerance = 1e-6;
    o->gradient_tolerance = 1e-10;
    o->parameter_tolerance = 1e-8;
    o->initial_trust_region_radius = 1e4;
    o->max_trust_region_radius = 1e16;
    o->min_trust_region_radius = 1e-32;
    o->min_relative_decrease = 1e-3;
    o->min_lm_diagonal = 1e-6;
    o->max_lm_diagonal = 1e32;
    o->max_num_consecutive_invalid_steps = 5;
}

typedef struct {
    int m;
    const double *coord;   /* 2m interleaved x,y */
    const double *flow;    /* 2m interleaved ux,uy */
    const double *alpha;   /* m */
    const double *alpha_k; /* m */
    int free_motion, free_k;
} LmProblem;

static inline void get_obs(const LmProblem *P, int i, OrcObs *o)
{
    o->x = P->coord[2 * i]; o->y = P->coord[2 * i + 1];
    o->ux = P->flow[2 * i]; o->uy = P->flow[2 * i + 1];
    o->alpha = P->alpha[i]; o->alpha_k = P->alpha_k[i];
}

/* Gather the free f-columns (Ceres order: v, w, k) out of the 7 analytic ones. */
static inline int pack_f(const LmProblem *P, double F[2][7], double Fp[2][7])
{
    int nf = 0;
    if (P->free_motion) {
        for (int j = 0; j < 6; ++j) { Fp[0][nf] = F[0][j]; Fp[1][nf] = F[1][j]; nf++; }
    }
    if (P->free_k) { Fp[0][nf] = F[0][6]; Fp[1][nf] = F[1][6]; nf++; }
    return nf;
}

static inline void unpack_params(const LmProblem *P, const double *f, const double *v0, const double *w0,
                                 double k0, double *v, double *w, double *k)
{
    int nf = 0;
    for (int j = 0; j < 3; ++j) { v[j] = v0[j]; w[j] = w0[j]; }
    *k = k0;
    if (P->free_motion) {
        for (int j = 0; j < 3; ++j) v[j] = f[nf++];
        for (int j = 0; j < 3; ++j) w[j] = f[nf++];
    }
    if (P->free_k) *k = f[nf++];
}

/* Evaluator::Evaluate.  cost = sum 0.5*|r_i|^2.  With jac: gradient of the f-block (gf),
 * Ceres' projected-gradient max norm |x - Plus(x,-g)|_inf over all free parameters (gmax) and,
 * if colsq_e/colsq_f are given, squared column norms of the unscaled Jacobian.
 * Returns 0 when a residual or a requested Jacobian entry is not finite (Ceres: evaluation
 * failure, residual_block.cc IsEvaluationValid). */
static int lm_evaluate(const LmProblem *P, const double *v, const double *w, double k, const double *d,
                       const double *fvec, int nf, double *r, int jac, double *cost_out, double *gmax_out,
                       double *colsq_e, double *colsq_f)
{
    const int m = P->m;
    double cost = 0.0, gmax = 0.0;
    double gf[7] = {0, 0, 0, 0, 0, 0, 0};
    double csf[7] = {0, 0, 0, 0, 0, 0, 0};
    int ok = 1;
    for (int i = 0; i < m; ++i) {
        OrcObs o;
        get_obs(P, i, &o);
        double ri[2];
        if (!jac) {
            rs_residual(&o, v, w, k, d[i], ri);
            if (!isfinite(ri[0]) || !isfinite(ri[1])) ok = 0;
        } else {
            double E[2], F[2][7], Fp[2][7];
            rs_residual_jac(&o, v, w, k, d[i], ri, E, F);
            pack_f(P, F, Fp);
            if (!isfinite(ri[0]) || !isfinite(ri[1]) || !isfinite(E[0]) || !isfinite(E[1])) ok = 0;
            for (int j = 0; j < nf; ++j) {
                if (!isfinite(Fp[0][j]) || !isfinite(Fp[1][j])) ok = 0;
                gf[j] += Fp[0][j] * ri[0] + Fp[1][j] * ri[1];
                csf[j] += Fp[0][j] * Fp[0][j] + Fp[1][j] * Fp[1][j];
            }
            double ge = E[0] * ri[0] + E[1] * ri[1];
            double proj = d[i] + (-ge);
            double diff = fabs(d[i] - proj);
            if (diff > gmax) gmax = diff;
            if (colsq_e) colsq_e[i] = E[0] * E[0] + E[1] * E[1];
        }
        if (r) { r[2 * i] = ri[0]; r[2 * i + 1] = ri[1]; }
        cost += 0.5 * (ri[0] * ri[0] + ri[1] * ri[1]);
    }
    if (jac) {
        for (int j = 0; j < nf; ++j) {
            double proj = fvec[j] + (-gf[j]);
            double diff = fabs(fvec[j] - proj);
            if (diff > gmax) gmax = diff;
            if (colsq_f) colsq_f[j] = csf[j];
        }
        if (gmax_out) *gmax_out = gmax;
    }
    *cost_out = cost;
    return ok;
}

static double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

/* Solves  min sum |r_i(v,w,k,d_i)|^2  the way ceres::Solve does for the reference's problems.
 * v,w,k,d are in/out.  Returns the termination type. */
static int orc_lm_solve(const LmProblem *P, const OrcLmOptions *opt, double *v_io, double *w_io,
                        double *k_io, double *d_io, OrcLmSummary *S)
{
    const int m = P->m;
    const int nf = (P->free_motion ? 6 : 0) + (P->free_k ? 1 : 0);
    OrcLmSummary local;
    if (!S) S = &local;
    memset(S, 0, sizeof *S);

    /* Problem / parameter validity (solver.cc: ParameterBlocksAreFinite) */
    {
        int finite = isfinite(*k_io);
        for (int j = 0; j < 3; ++j) finite = finite && isfinite(v_io[j]) && isfinite(w_io[j]);
        for (int i = 0; i < m && finite; ++i) finite = finite && isfinite(d_io[i]);
        if (!finite) {
            S->termination = ORC_FAILURE; S->reason = ORC_REASON_NONFINITE_INPUT;
            return S->termination;
        }
    }
    if (m == 0) { /* no residual blocks: Ceres returns CONVERGENCE without touching anything */
        S->termination = ORC_CONVERGENCE; S->reason = ORC_REASON_FUNCTION_TOL;
        return S->termination;
    }

    double *r = (double *)malloc(sizeof(double) * 2 * m);
    double *d = (double *)malloc(sizeof(double) * m);        /* x (depth part) */
    double *d_cand = (double *)malloc(sizeof(double) * m);
    double *scale_e = (double *)malloc(sizeof(double) * m);
    double *diag_e = (double *)malloc(sizeof(double) * m);
    double f[7], f_cand[7], scale_f[7], diag_f[7];
    double v[3], w[3], k;
    memcpy(d, d_io, sizeof(double) * m);
    {
        int q = 0;
        if (P->free_motion) { for (int j = 0; j < 3; ++j) f[q++] = v_io[j]; for (int j = 0; j < 3; ++j) f[q++] = w_io[j]; }
        if (P->free_k) f[q++] = *k_io;
    }
    unpack_params(P, f, v_io, w_io, *k_io, v, w, &k);

    double x_cost = 0.0, gmax = 0.0, x_norm = 0.0;
    int termination = ORC_NO_CONVERGENCE, reason = ORC_REASON_NONE;
    int usable = 1;
    double radius = opt->initial_trust_region_radius;
    double decrease_factor = 2.0;
    int reuse_diagonal = 0;
    int iteration = 0;
    int step_is_successful = 0;
    int num_consecutive_invalid = 0;

    /* ---- IterationZero ---- */
    if (!lm_evaluate(P, v, w, k, d, f, nf, r, 1, &x_cost, &gmax, scale_e, scale_f)) {
        termination = ORC_FAILURE; reason = ORC_REASON_EVAL_FAILED; usable = 0;
        goto done;
    }
    for (int i = 0; i < m; ++i) scale_e[i] = 1.0 / (1.0 + sqrt(scale_e[i]));
    for (int j = 0; j < nf; ++j) scale_f[j] = 1.0 / (1.0 + sqrt(scale_f[j]));
    {
        double s = 0.0;
        for (int i = 0; i < m; ++i) s += d[i] * d[i];
        for (int j = 0; j < nf; ++j) s += f[j] * f[j];
        x_norm = sqrt(s);
    }
    S->initial_cost = x_cost;
    S->trace_cost[0] = x_cost;
    S->trace_radius[0] = radius;
    /* IterationZero ends with step_is_valid = step_is_successful = true (trust_region_minimizer.cc): iteration 0
     * is counted among the successful steps and the gradient tolerance IS tested before the first step */
    step_is_successful = 1;

    for (;;) {
        /* ---- FinalizeIterationAndCheckIfMinimizerCanContinue ---- */
        if (step_is_successful) S->num_successful++; else S->num_unsuccessful++;
        if (iteration >= opt->max_num_iterations) { termination = ORC_NO_CONVERGENCE; reason = ORC_REASON_MAX_ITER; break; }
        if (step_is_successful && gmax <= opt->gradient_tolerance) { termination = ORC_CONVERGENCE; reason = ORC_REASON_GRADIENT_TOL; break; }
        if (radius <= opt->min_trust_region_radius) { termination = ORC_CONVERGENCE; reason = ORC_REASON_MIN_RADIUS; break; }
        iteration++;
        step_is_successful = 0;

        /* ---- ComputeTrustRegionStep: LevenbergMarquardtStrategy::ComputeStep ---- */
        double lhs[49], rhs[7], yf[7], step_f[7];
        for (int a = 0; a < nf * nf; ++a) lhs[a] = 0.0;
        for (int a = 0; a < nf; ++a) rhs[a] = 0.0;
        double diag_f_new[7] = {0, 0, 0, 0, 0, 0, 0};
        /* pass A: diagonal (when not reused) + SchurEliminator::Eliminate */
        for (int i = 0; i < m; ++i) {
            OrcObs o; get_obs(P, i, &o);
            double ri[2], E[2], F[2][7], Fp[2][7];
            rs_residual_jac(&o, v, w, k, d[i], ri, E, F);
            pack_f(P, F, Fp);
            ri[0] = r[2 * i]; ri[1] = r[2 * i + 1];
            double e0 = E[0] * scale_e[i], e1 = E[1] * scale_e[i];
            for (int j = 0; j < nf; ++j) { Fp[0][j] *= scale_f[j]; Fp[1][j] *= scale_f[j]; }
            double ee = e0 * e0 + e1 * e1;
            if (!reuse_diagonal) {
                diag_e[i] = clampd(ee, opt->min_lm_diagonal, opt->max_lm_diagonal);
                for (int j = 0; j < nf; ++j) diag_f_new[j] += Fp[0][j] * Fp[0][j] + Fp[1][j] * Fp[1][j];
            }
            if (nf > 0) {
                double De = sqrt(diag_e[i] / radius);
                double ete = ee + De * De;
                double inv = 1.0 / ete;
                double ge = e0 * ri[0] + e1 * ri[1];
                double t = inv * ge;
                double sj0 = ri[0] - e0 * t, sj1 = ri[1] - e1 * t;
                double buf[7];
                for (int j = 0; j < nf; ++j) buf[j] = e0 * Fp[0][j] + e1 * Fp[1][j];
                for (int j = 0; j < nf; ++j) {
                    rhs[j] += Fp[0][j] * sj0 + Fp[1][j] * sj1;
                    double bj = buf[j] * inv;
                    for (int c = j; c < nf; ++c)
                        lhs[j * nf + c] += Fp[0][j] * Fp[0][c] + Fp[1][j] * Fp[1][c] - bj * buf[c];
                }
            }
        }
        if (!reuse_diagonal)
            for (int j = 0; j < nf; ++j) diag_f[j] = clampd(diag_f_new[j], opt->min_lm_diagonal, opt->max_lm_diagonal);
        int solver_ok = 1;
        if (nf > 0) {
            for (int j = 0; j < nf; ++j) {
                double Df = sqrt(diag_f[j] / radius);
                lhs[j * nf + j] += Df * Df;
                for (int c = 0; c < j; ++c) lhs[j * nf + c] = lhs[c * nf + j];
            }
            if (orc_cholesky_solve(lhs, nf, rhs, yf)) solver_ok = 0;
        }
        reuse_diagonal = 1;
        /* pass B: BackSubstitute, model cost change, candidate point */
        double model_cost_change = 0.0, step_sq = 0.0;
        int step_finite = solver_ok;
        if (solver_ok) {
            for (int j = 0; j < nf; ++j) { step_f[j] = -yf[j]; if (!isfinite(yf[j])) step_finite = 0; }
            double mcc = 0.0;
            for (int i = 0; i < m; ++i) {
                OrcObs o; get_obs(P, i, &o);
                double ri[2], E[2], F[2][7], Fp[2][7];
                rs_residual_jac(&o, v, w, k, d[i], ri, E, F);
                pack_f(P, F, Fp);
                ri[0] = r[2 * i]; ri[1] = r[2 * i + 1];
                double e0 = E[0] * scale_e[i], e1 = E[1] * scale_e[i];
                for (int j = 0; j < nf; ++j) { Fp[0][j] *= scale_f[j]; Fp[1][j] *= scale_f[j]; }
                double ee = e0 * e0 + e1 * e1;
                double De = sqrt(diag_e[i] / radius);
                double ete = ee + De * De;
                double inv = 1.0 / ete;
                double sj0 = ri[0], sj1 = ri[1];
                for (int j = 0; j < nf; ++j) { sj0 -= Fp[0][j] * yf[j]; sj1 -= Fp[1][j] * yf[j]; }
                double ye = (e0 * sj0 + e1 * sj1) * inv;
                if (!isfinite(ye)) step_finite = 0;
                double step_e = -ye;
                double mr0 = e0 * step_e, mr1 = e1 * step_e;
                for (int j = 0; j < nf; ++j) { mr0 += Fp[0][j] * step_f[j]; mr1 += Fp[1][j] * step_f[j]; }
                mcc += mr0 * (ri[0] + mr0 / 2.0) + mr1 * (ri[1] + mr1 / 2.0);
                d_cand[i] = d[i] + step_e * scale_e[i];
                double dd = d[i] - d_cand[i];
                step_sq += dd * dd;
            }
            model_cost_change = -mcc;
        }
        int step_is_valid = step_finite && (model_cost_change > 0.0);
        if (!step_is_valid) {
            /* HandleInvalidStep */
            num_consecutive_invalid++;
            if (num_consecutive_invalid >= opt->max_num_consecutive_invalid_steps) {
                termination = ORC_FAILURE; reason = ORC_REASON_INVALID_STEPS; usable = 0;
                S->num_unsuccessful++;
                break;
            }
            radius = radius / decrease_factor;   /* StepIsInvalid -> StepRejected(0) */
            decrease_factor *= 2.0;
            reuse_diagonal = 1;
            if (iteration < 64) { S->trace_cost[iteration] = x_cost; S->trace_radius[iteration] = radius; S->trace_rho[iteration] = 0.0; S->trace_accepted[iteration] = -1; }
            continue;
        }
        num_consecutive_invalid = 0;
        for (int j = 0; j < nf; ++j) {
            f_cand[j] = f[j] + step_f[j] * scale_f[j];
            double dd = f[j] - f_cand[j];
            step_sq += dd * dd;
        }
        /* ComputeCandidatePointAndEvaluateCost */
        double vc[3], wc[3], kc, cand_cost;
        unpack_params(P, f_cand, v_io, w_io, *k_io, vc, wc, &kc);
        if (!lm_evaluate(P, vc, wc, kc, d_cand, f_cand, nf, NULL, 0, &cand_cost, NULL, NULL, NULL))
            cand_cost = DBL_MAX;
        /* ParameterToleranceReached */
        double step_norm = sqrt(step_sq);
        if (step_norm <= opt->parameter_tolerance * (x_norm + opt->parameter_tolerance)) {
            termination = ORC_CONVERGENCE; reason = ORC_REASON_PARAMETER_TOL;
            break;   /* returns x, the candidate is discarded */
        }
        /* FunctionToleranceReached */
        double cost_change = x_cost - cand_cost;
        if (fabs(cost_change) <= opt->function_tolerance * x_cost) {
            termination = ORC_CONVERGENCE; reason = ORC_REASON_FUNCTION_TOL;
            break;   /* returns x, the candidate is discarded */
        }
        /* IsStepSuccessful (monotonic steps: step quality = relative decrease) */
        double rho = cost_change / model_cost_change;
        if (rho > opt->min_relative_decrease) {
            /* HandleSuccessfulStep */
            memcpy(d, d_cand, sizeof(double) * m);
            for (int j = 0; j < nf; ++j) f[j] = f_cand[j];
            unpack_params(P, f, v_io, w_io, *k_io, v, w, &k);
            {
                double s = 0.0;
                for (int i = 0; i < m; ++i) s += d[i] * d[i];
                for (int j = 0; j < nf; ++j) s += f[j] * f[j];
                x_norm = sqrt(s);
            }
            if (!lm_evaluate(P, v, w, k, d, f, nf, r, 1, &x_cost, &gmax, NULL, NULL)) {
                termination = ORC_FAILURE; reason = ORC_REASON_EVAL_FAILED; usable = 0;
                break;
            }
            step_is_successful = 1;
            radius = radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * rho - 1.0, 3));
            radius = fmin(opt->max_trust_region_radius, radius);
            decrease_factor = 2.0;
            reuse_diagonal = 0;
        } else {
            /* HandleUnsuccessfulStep */
            radius = radius / decrease_factor;
            decrease_factor *= 2.0;
            reuse_diagonal = 1;
        }
        if (iteration < 64) {
            S->trace_cost[iteration] = step_is_successful ? x_cost : cand_cost;
            S->trace_radius[iteration] = radius;
            S->trace_rho[iteration] = rho;
            S->trace_accepted[iteration] = step_is_successful;
        }
    }

done:
    S->termination = termination;
    S->reason = reason;
    S->iterations = iteration;   /* number of trust-region steps computed (incl. a final discarded one) */
    S->final_radius = radius;
    S->final_gradient_max_norm = gmax;
    if (usable) {
        /* solver.cc Minimize(): the last accepted state is written back */
        S->final_cost = x_cost;
        memcpy(d_io, d, sizeof(double) * m);
        unpack_params(P, f, v_io, w_io, *k_io, v, w, &k);
        for (int j = 0; j < 3; ++j) { v_io[j] = v[j]; w_io[j] = w[j]; }
        *k_io = k;
    } else {
        S->final_cost = S->initial_cost; /* FAILURE: original parameters restored */
    }
    free(r); free(d); free(d_cand); free(scale_e); free(diag_e);
    return termination;
}

/* a8: nonlinear_refinement::estimateInverseDepths          nonlinearRefinement.cc:109-180 */
ORC_API int orc_estimate_inverse_depths(const double *coord, const double *flow, int n, const double *v,
                                        const double *w, double k, const double *alpha,
                                        const double *alpha_k, double *inv_depth, OrcLmSummary *summary)
{
    LmProblem P = {n, coord, flow, alpha, alpha_k, 0, 0};
    OrcLmOptions opt;
    orc_lm_default_options(&opt);
    double vv[3] = {v[0], v[1], v[2]}, ww[3] = {w[0], w[1], w[2]}, kk = k;
    for (int i = 0; i < n; ++i) inv_depth[i] = 1.0;           /* :140 */
    return orc_lm_solve(&P, &opt, vv, ww, &kk, inv_depth, summary);
}

/* single pixel variant (dead code in the reference)          nonlinearRefinement.cc:55-106 */
ORC_API double orc_estimate_inverse_depth(const double *coord2, const double *v, const double *w,
                                          const double *flow2, double k, double alpha, double alpha_k)
{
    double d = 1.0;
    orc_estimate_inverse_depths(coord2, flow2, 1, v, w, k, &alpha, &alpha_k, &d, NULL);
    return d;
}

/* ------------------------------------------------------------------------------------------ */
/* a6: minimal::ransac with an INJECTED hypothesis / sample list    minimal.cc:209-306         */
/* ------------------------------------------------------------------------------------------ */
/* Scoring of one hypothesis (w,v,k): depth estimation over all n points, then the inlier loop
 * :255-275.  mask (n bytes) and inv_depth (n) are outputs. */
ORC_API int orc_score_hypothesis(const double *q, const double *u, const double *alpha, const double *alpha_k,
                                 int n, const double *hyp7, double tolerance, uint8_t *mask,
                                 double *inv_depth, double *inlier_error_out, OrcLmSummary *summary)
{
    const double *w = hyp7, *v = hyp7 + 3;
    double k = hyp7[6];
    orc_estimate_inverse_depths(q, u, n, v, w, k, alpha, alpha_k, inv_depth, summary);
    int num_inliers = 0;
    double inlier_error = 0;
    for (int j = 0; j < n; ++j) {
        double x = q[2 * j], y = q[2 * j + 1];
        /* A = [1 0 -x; 0 1 -y], B = [-xy (1+x^2) -y; -(1+y^2) xy x]; products written out
         * left-to-right, including the structural 1* and 0* terms (NaN/Inf propagate alike) */
        double av0 = 1.0 * v[0] + 0.0 * v[1] + (-x) * v[2];
        double av1 = 0.0 * v[0] + 1.0 * v[1] + (-y) * v[2];
        double bw0 = (-x * y) * w[0] + (1 + x * x) * w[1] + (-y) * w[2];
        double bw1 = (-(1 + y * y)) * w[0] + (x * y) * w[1] + x * w[2];
        double beta = (alpha[j] + k * alpha_k[j]) * (2.0 / (2.0 + k));
        double ue0 = beta * (av0 * inv_depth[j] + bw0);
        double ue1 = beta * (av1 * inv_depth[j] + bw1);
        double dx = ue0 - u[2 * j], dy = ue1 - u[2 * j + 1];
        double error = sqrt(dx * dx + dy * dy);
        int in = (error < tolerance);
        mask[j] = (uint8_t)in;
        if (in) { num_inliers++; inlier_error += error; }
    }
    *inlier_error_out = inlier_error;
    return num_inliers;
}

/* The RANSAC loop with the random draw replaced by a caller-supplied list.
 *   mode 0: hyps = H x 7 (w,v,k) hypotheses, scored as given
 *   mode 1: samples = H x 9 point indices; the 9-point solver is run on them (:230-247)
 * Outputs: counts[H], sumerr[H], best_idx, best hypothesis (7), mask_best[n], inv_depth_best[n].
 * Returns num_inliers_best. */
ORC_API int orc_ransac(const double *q, const double *u, const double *alpha, const double *alpha_k, int n,
                       int use_alpha_k, int mode, const double *hyps, const int *samples, int H,
                       double tolerance, int *counts, double *sumerr, int *best_idx, double *best7,
                       uint8_t *mask_best, double *inv_depth_best, double *hyps_out)
{
    uint8_t *mask = (uint8_t *)malloc(n > 0 ? n : 1);
    double *inv_depth = (double *)malloc(sizeof(double) * (n > 0 ? n : 1));
    int num_inliers_best = -1;
    double inlier_error_best = 0;
    for (int j = 0; j < 7; ++j) best7[j] = 0.0;
    *best_idx = -1;
    for (int i = 0; i < n; ++i) { inv_depth_best[i] = 0.0; mask_best[i] = 0; }
    for (int it = 0; it < H; ++it) {
        double hyp[7];
        if (mode == 0) {
            memcpy(hyp, hyps + 7 * it, sizeof hyp);
        } else {
            double cq[18], cu[18], ca[9], cak[9];
            for (int j = 0; j < 9; ++j) {
                int index = samples[9 * it + j];
                cq[2 * j] = q[2 * index]; cq[2 * j + 1] = q[2 * index + 1];
                cu[2 * j] = u[2 * index]; cu[2 * j + 1] = u[2 * index + 1];
                ca[j] = alpha[index]; cak[j] = alpha_k[index];
            }
            orc_calculate_velocities(cq, cu, ca, cak, use_alpha_k, hyp);
        }
        if (hyps_out) memcpy(hyps_out + 7 * it, hyp, sizeof hyp);
        double inlier_error;
        int num_inliers = orc_score_hypothesis(q, u, alpha, alpha_k, n, hyp, tolerance, mask, inv_depth,
                                               &inlier_error, NULL);
        if (counts) counts[it] = num_inliers;
        if (sumerr) sumerr[it] = inlier_error;
        if (num_inliers > num_inliers_best ||
            (num_inliers == num_inliers_best && inlier_error < inlier_error_best)) {
            num_inliers_best = num_inliers;
            memcpy(mask_best, mask, n);
            memcpy(best7, hyp, sizeof hyp);
            inlier_error_best = inlier_error;
            memcpy(inv_depth_best, inv_depth, sizeof(double) * n);
            *best_idx = it;
        }
    }
    free(mask); free(inv_depth);
    return num_inliers_best;
}

/* tail of ransac(): gather the consensus set in ascending index order      minimal.cc:291-305 */
ORC_API int orc_gather_inliers(const double *q, const double *alpha, const double *alpha_k, int n,
                               const uint8_t *mask, const double *inv_depth, double *inliers3,
                               double *alpha_in, double *alpha_k_in)
{
    int j = 0;
    for (int i = 0; i < n; ++i) {
        if (mask[i]) {
            inliers3[3 * j] = q[2 * i];
            inliers3[3 * j + 1] = q[2 * i + 1];
            inliers3[3 * j + 2] = 1.0 / inv_depth[i];
            alpha_in[j] = alpha[i];
            alpha_k_in[j] = alpha_k[i];
            j++;
        }
    }
    return j;
}

/* ------------------------------------------------------------------------------------------ */
/* a9: nonlinear_refinement::nonLinearRefinement          nonlinearRefinement.cc:183-252       */
/* ------------------------------------------------------------------------------------------ */
/* flow: the array the caller passed (Q1: residual i uses flow(:,i) of THIS array, which the
 * reference's callers leave un-gathered).  inliers3: 3 x m (x, y, z=1/d).  v,w,k in/out.
 * z_out[m] = refined depth (1/d).  fixed_pairing != 0 selects the repaired pairing instead of
 * the reference behaviour: flow is then indexed through flow_index[m]. */
ORC_API int orc_nonlinear_refinement(const double *flow, const double *inliers3, const double *alpha,
                                     const double *alpha_k, int m, double *v, double *w, double *k,
                                     int const_acceleration, const int *flow_index, double *z_out,
                                     OrcLmSummary *summary)
{
    double *coord = (double *)malloc(sizeof(double) * 2 * (m > 0 ? m : 1));
    double *fl = (double *)malloc(sizeof(double) * 2 * (m > 0 ? m : 1));
    double *d = (double *)malloc(sizeof(double) * (m > 0 ? m : 1));
    for (int i = 0; i < m; ++i) {
        coord[2 * i] = inliers3[3 * i];
        coord[2 * i + 1] = inliers3[3 * i + 1];
        int fi = flow_index ? flow_index[i] : i;
        fl[2 * i] = flow[2 * fi];
        fl[2 * i + 1] = flow[2 * fi + 1];
        d[i] = 1.0 / inliers3[3 * i + 2];
    }
    LmProblem P = {m, coord, fl, alpha, alpha_k, 1, const_acceleration ? 1 : 0};
    OrcLmOptions opt;
    orc_lm_default_options(&opt);
    int term = orc_lm_solve(&P, &opt, v, w, k, d, summary);
    for (int i = 0; i < m; ++i) z_out[i] = 1.0 / d[i];
    free(coord); free(fl); free(d);
    return term;
}

/* ------------------------------------------------------------------------------------------ */
/* a10: sign fix + depth raster glue          main.cc:466-509, errorMeasure.cpp:162-210        */
/* ------------------------------------------------------------------------------------------ */
/* inliers3 (x,y,z) and v are modified in place (sign fix).  depth_map: rows*cols COLUMN-major
 * (Eigen MatrixXd), zero-filled here.  depth_img (may be NULL): rows*cols row-major 8UC1.
 * z_min_init: INFINITY (main.cc:482) or 100000 (errorMeasure.cpp:188).
 * Out-of-image raster targets are undefined behaviour in the reference (Q3); they are skipped. */
ORC_API void orc_depth_glue(double *inliers3, int m, double *v, double fx, double fy, double cx, double cy,
                            int rows, int cols, double z_min_init, double *depth_map, uint8_t *depth_img,
                            double *zmean_out)
{
    double count_z = 0;
    for (int i = 0; i < m; ++i) count_z += inliers3[3 * i + 2];
    double z_mean = count_z * 1.0 / m;
    if (zmean_out) *zmean_out = z_mean;
    if (z_mean < 0) {
        for (int i = 0; i < m; ++i) inliers3[3 * i + 2] *= -1.0;
        for (int j = 0; j < 3; ++j) v[j] *= -1.0;
    }
    double z_min = z_min_init, z_max = 0;
    for (int i = 0; i < m; ++i) {
        if (inliers3[3 * i + 2] < z_min) z_min = inliers3[3 * i + 2];
        if (inliers3[3 * i + 2] > z_max) z_max = inliers3[3 * i + 2];
    }
    const int min_z_value = 10;
    double multiplier = 244.0 / (z_max - z_min);
    memset(depth_map, 0, sizeof(double) * (size_t)rows * cols);
    if (depth_img) memset(depth_img, 0, (size_t)rows * cols);
    for (int i = 0; i < m; ++i) {
        double xd = fx * inliers3[3 * i] + cx + 0.5;
        double yd = fy * inliers3[3 * i + 1] + cy + 0.5;
        if (!(fabs(xd) < 2147483648.0) || !(fabs(yd) < 2147483648.0)) continue;
        int x = (int)xd, y = (int)yd;
        if (x < 0 || x >= cols || y < 0 || y >= rows) continue;
        if (depth_img) {
            double zz = (inliers3[3 * i + 2] - z_min) * multiplier;
            int z = min_z_value;
            if (fabs(zz) < 2147483000.0) z = min_z_value + (int)zz;
            depth_img[(size_t)y * cols + x] = (uint8_t)z;
        }
        depth_map[(size_t)y + (size_t)x * rows] = inliers3[3 * i + 2];
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a12: RsFrame::setRelativePose                                   rsframe.cc:771-800          */
/* ------------------------------------------------------------------------------------------ */
/* R: rows x 9 row-major 3x3, t: rows x 3 */
ORC_API void orc_set_relative_pose(const double *v, const double *w, double k, double gamma, int rows,
                                   double *R, double *t)
{
    static const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double skew[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    if (rows <= 0) return;
    memcpy(R, I3, sizeof I3);
    t[0] = t[1] = t[2] = 0.0;
    for (int i = 1; i < rows; ++i) {
        double beta_1 = (gamma * i / rows + 0.5 * k * (gamma * gamma * i * i) / (rows * rows)) * (2.0 / (2.0 + k));
        double Rn[9];
        for (int a = 0; a < 9; ++a) Rn[a] = I3[a] + beta_1 * skew[a];
        mat3_mul(I3, Rn, R + 9 * i);                         /* R0 * R_new, R0 = I */
        for (int a = 0; a < 3; ++a) t[3 * i + a] = 0.0 + beta_1 * v[a];
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a13 + a14: planeToSpace / cameraToWorldFrame / worldToCameraFrame / spaceToPlane and         */
/* RsFrame::backProject(Gs)                        rsframe.cc:629-736, 803-878                  */
/* ------------------------------------------------------------------------------------------ */
/* image: rows*cols*3 BGR row-major.  depth_map: COLUMN-major rows x cols.  K4 = fx,fy,cx,cy.
 * gs_mode: 0 = backProject (pose of scanline y), 1 = backProjectGs (pose of scanline 0).
 * gs_out: rows*cols*3, zero-filled then splatted in raster order (last writer wins).
 * coords3d (may be NULL): rows*cols*3 float; entries of skipped pixels are uninitialised in
 * the reference (Q16) and written as 0 here. */
static inline int trunc_to_int(double a, int *out)
{
    if (!(fabs(a) < 2147483648.0)) return 0;  /* NaN/Inf/overflow: x86 cvttsd2si gives INT_MIN */
    *out = (int)a;
    return 1;
}

ORC_API void orc_back_project(const uint8_t *image, const double *depth_map, int rows, int cols,
                              const double *K4, const double *R, const double *t, int gs_mode,
                              uint8_t *gs_out, float *coords3d)
{
    const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
    memset(gs_out, 0, (size_t)rows * cols * 3);
    if (coords3d) memset(coords3d, 0, sizeof(float) * (size_t)rows * cols * 3);
    for (int y = 0; y < rows; ++y) {
        const int s = gs_mode ? 0 : y;
        const double *Rs = R + 9 * s, *ts = t + 3 * s;
        const double *R0 = R, *t0 = t;
        /* inverse pose of scanline s: [R^T | -R^T t]  (:719-733) */
        double Rt[9], ti[3];
        mat3_t(Rs, Rt);
        for (int a = 0; a < 3; ++a)
            ti[a] = (-Rt[a * 3 + 0]) * ts[0] + (-Rt[a * 3 + 1]) * ts[1] + (-Rt[a * 3 + 2]) * ts[2];
        for (int x = 0; x < cols; ++x) {
            const uint8_t *px = image + ((size_t)y * cols + x) * 3;
            if (px[0] == 1 && px[1] == 1 && px[2] == 1) continue;             /* :815 */
            /* planeToSpace(Vector2d(x,y)) :646-665 */
            double nx = ((double)x - cx) * 1.0 / fx;
            double ny = ((double)y - cy) * 1.0 / fy;
            double z = depth_map[(size_t)y + (size_t)x * rows];
            double Pc[3] = {z * nx, z * ny, z * 1.0};
            /* cameraToWorldFrame :712-736 (4x4 times homogeneous point) */
            double Pw[3];
            for (int a = 0; a < 3; ++a)
                Pw[a] = Rt[a * 3 + 0] * Pc[0] + Rt[a * 3 + 1] * Pc[1] + Rt[a * 3 + 2] * Pc[2] + ti[a] * 1.0;
            /* worldToCameraFrame(., 0) :687-708 */
            double Pg[3];
            for (int a = 0; a < 3; ++a)
                Pg[a] = R0[a * 3 + 0] * Pw[0] + R0[a * 3 + 1] * Pw[1] + R0[a * 3 + 2] * Pw[2] + t0[a] * 1.0;
            /* spaceToPlane :629-642 -- y uses f_x too (Q12) */
            double u = Pg[0] / Pg[2] * fx + cx;
            double vv = Pg[1] / Pg[2] * fx + cy;
            if (coords3d) {
                float *c = coords3d + ((size_t)y * cols + x) * 3;
                c[0] = (float)Pw[0]; c[1] = (float)Pw[1]; c[2] = (float)Pw[2];
            }
            int tx, ty;
            if (!trunc_to_int(u + 0.5, &tx) || !trunc_to_int(vv + 0.5, &ty)) continue;
            if (tx >= 0 && tx < cols && ty >= 0 && ty < rows) {
                uint8_t *o = gs_out + ((size_t)ty * cols + tx) * 3;
                o[0] = px[0]; o[1] = px[1]; o[2] = px[2];
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* a15: Camera::interpolateCrackyImage                             camera.cc:694-774           */
/* ------------------------------------------------------------------------------------------ */
static inline int is_black(const uint8_t *p)
{   /* cv::norm(Vec3b) <= 15  <=>  b^2+g^2+r^2 <= 225 (sqrt is monotone and sqrt(225)=15 exactly) */
    int s = (int)p[0] * p[0] + (int)p[1] * p[1] + (int)p[2] * p[2];
    return s <= 225;
}
static inline uint8_t saturate_u8(double v)
{   /* cv::saturate_cast<uchar>(double): cvRound (round half to even) then clamp */
    long iv = lrint(v);
    return (uint8_t)(iv < 0 ? 0 : (iv > 255 ? 255 : iv));
}

ORC_API void orc_interpolate_cracky_image(const uint8_t *in, int rows, int cols, unsigned offset,
                                          uint8_t *out)
{
    memcpy(out, in, (size_t)rows * cols * 3);
    const int off = (int)offset;
    for (int row = off; row < rows - off; ++row) {
        for (int col = off; col < cols - off; ++col) {
            const uint8_t *poi = in + ((size_t)row * cols + col) * 3;
            if (!is_black(poi)) continue;
            const uint8_t *nb[4] = {
                in + ((size_t)(row - off) * cols + col) * 3,   /* above */
                in + ((size_t)(row + off) * cols + col) * 3,   /* below */
                in + ((size_t)row * cols + (col - off)) * 3,   /* left  */
                in + ((size_t)row * cols + (col + off)) * 3};  /* right */
            double sum[3] = {0, 0, 0};
            unsigned count = 0;
            for (int a = 0; a < 4; ++a) {
                if (!is_black(nb[a])) {
                    sum[0] += nb[a][0]; sum[1] += nb[a][1]; sum[2] += nb[a][2];
                    count++;
                }
            }
            if (count > 0) {
                uint8_t *o = out + ((size_t)row * cols + col) * 3;
                double f = 1 / (double)count;
                o[0] = saturate_u8(f * sum[0]);
                o[1] = saturate_u8(f * sum[1]);
                o[2] = saturate_u8(f * sum[2]);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Driver for the timed region "refine + rectify" of one frame pair (main.cc:457-523)          */
/* ------------------------------------------------------------------------------------------ */
/* Inputs: normalised flow (2 x >=m, un-gathered, Q1), RANSAC result (inliers3 3 x m with z,
 * alpha[m], alpha_k[m], v, w, k), RS image, K, gamma.  Outputs: refined v,w,k; z_out[m];
 * depth_map (column-major); gs image after crack fill.  Scratch is allocated internally. */
ORC_API int orc_refine_rectify(const double *flow, const double *inliers3_in, const double *alpha,
                               const double *alpha_k, int m, double *v, double *w, double *k,
                               int const_acceleration, int gs_mode, const uint8_t *image, int rows,
                               int cols, const double *K4, double gamma, double *z_out,
                               double *depth_map, uint8_t *rectified, OrcLmSummary *summary)
{
    double *inl = (double *)malloc(sizeof(double) * 3 * (m > 0 ? m : 1));
    memcpy(inl, inliers3_in, sizeof(double) * 3 * m);
    int term = orc_nonlinear_refinement(flow, inl, alpha, alpha_k, m, v, w, k, const_acceleration, NULL,
                                        z_out, summary);
    for (int i = 0; i < m; ++i) inl[3 * i + 2] = z_out[i];
    orc_depth_glue(inl, m, v, K4[0], K4[1], K4[2], K4[3], rows, cols, INFINITY, depth_map, NULL, NULL);
    for (int i = 0; i < m; ++i) z_out[i] = inl[3 * i + 2];
    double *R = (double *)malloc(sizeof(double) * 9 * rows);
    double *t = (double *)malloc(sizeof(double) * 3 * rows);
    uint8_t *gs = (uint8_t *)malloc((size_t)rows * cols * 3);
    orc_set_relative_pose(v, w, *k, gamma, rows, R, t);
    orc_back_project(image, depth_map, rows, cols, K4, R, t, gs_mode, gs, NULL);
    orc_interpolate_cracky_image(gs, rows, cols, 1, rectified);
    free(inl); free(R); free(t); free(gs);
    return term;
}

ORC_API int orc_sizeof_summary(void) { return (int)sizeof(OrcLmSummary); }

/* ------------------------------------------------------------------------------------------ */
/* Camera::meanReprojectionError / createErrorImage   camera.cc:503-691         */
/*                RsFrame::getGroundtruthDepthMap rsframe.cc:416-436, relocatePose :953-967    */
/* ------------------------------------------------------------------------------------------ */
/* Eigen 3.3 Inverse.h, compute_inverse_size3_helper: cofactors of column 0, determinant as their
 * dot product with column 0, every entry = cofactor * (1/det). */
static double cof3(const double *m, int i, int j)
{
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
}
static void mat3_inverse_eigen(const double *m, double *inv)
{
    const double c00 = cof3(m, 0, 0), c10 = cof3(m, 1, 0), c20 = cof3(m, 2, 0);
    const double det = c00 * m[0] + c10 * m[3] + c20 * m[6];
    const double invdet = 1.0 / det;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) inv[r * 3 + c] = cof3(m, c, r) * invdet;      /* inverse(r,c) = cofactor<c,r> / det */
}

/* unprojection maps ux,uy,uz and the output are COLUMN-major rows x cols (Eigen MatrixXd);
 * R,t: ground-truth scanline poses, rows x 9 (row-major 3x3) and rows x 3. */
ORC_API void orc_groundtruth_depth_map(const double *ux, const double *uy, const double *uz, const double *R,
                                       const double *t, int rows, int cols, double *depth)
{
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            const size_t i = (size_t)y + (size_t)x * rows;
            const double P[3] = {ux[i], uy[i], uz[i]};
            depth[i] = 0.0;
            if (sqrt(P[0] * P[0] + P[1] * P[1] + P[2] * P[2]) > 0) {                  /* :428 */
                const double *Rs = R + 9 * y, *ts = t + 3 * y;                          /* worldToCameraFrame :687-708 */
                depth[i] = Rs[6] * P[0] + Rs[7] * P[1] + Rs[8] * P[2] + ts[2] * 1.0;
            }
        }
}

ORC_API void orc_relocate_pose(double *R, double *t, int rows)
{
    double R0[9], R0inv[9], t0[3];
    memcpy(R0, R, sizeof R0);
    memcpy(t0, t, sizeof t0);
    mat3_inverse_eigen(R0, R0inv);                                   /* evaluated per row in the reference: same value */
    for (int i = 1; i < rows; ++i) {                                  /* row 0 keeps its pose (:961) */
        double Rn[9];
        for (int a = 0; a < 3; ++a) t[3 * i + a] = t[3 * i + a] - t0[a];
        mat3_mul(R0inv, R + 9 * i, Rn);
        memcpy(R + 9 * i, Rn, sizeof Rn);
    }
}

/* coords3d_est: rows*cols*3 float (RsFrame::get3dCoordinates after backProject); depth_est: the
 * frame's depth_map_ (column-major), used by planeToSpace when the ground-truth depth is exactly 0
 * (rsframe.cc:657-659).  Returns the mean error; error_image (nullable): rows*cols, row-major. */
ORC_API double orc_mean_reprojection_error(const float *coords3d_est, const double *ux, const double *uy,
                                           const double *uz, const double *R_gt, const double *t_gt,
                                           const double *depth_est, int rows, int cols, const double *K4,
                                           double max_norm, double *mean_scale, int *num_outliers,
                                           int *points_used, uint8_t *error_image)
{
    const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
    const size_t tot = (size_t)rows * cols;
    double *gt_depth = (double *)malloc(sizeof(double) * tot);
    double *R = (double *)malloc(sizeof(double) * 9 * rows), *t = (double *)malloc(sizeof(double) * 3 * rows);
    float *truep = (float *)malloc(sizeof(float) * 3 * tot);
    double *scales = (double *)calloc(3 * tot, sizeof(double));
    orc_groundtruth_depth_map(ux, uy, uz, R_gt, t_gt, rows, cols, gt_depth);       /* camera.cc:612, original poses */
    memcpy(R, R_gt, sizeof(double) * 9 * rows);
    memcpy(t, t_gt, sizeof(double) * 3 * rows);
    orc_relocate_pose(R, t, rows);                                                 /* :620 */
    int outliers = 0;
    for (int x = 0; x < cols; ++x)
        for (int y = 0; y < rows; ++y) {
            double z = gt_depth[(size_t)y + (size_t)x * rows];
            const double nx = ((double)x - cx) * 1.0 / fx, ny = ((double)y - cy) * 1.0 / fy;
            if (z == 0) z = depth_est[(size_t)y + (size_t)x * rows];
            const double Pc[3] = {z * nx, z * ny, z * 1.0};
            const double *Rs = R + 9 * y, *ts = t + 3 * y;
            double Rt[9], ti[3], Pw[3];
            mat3_t(Rs, Rt);
            for (int a = 0; a < 3; ++a)
                ti[a] = (-Rt[a * 3 + 0]) * ts[0] + (-Rt[a * 3 + 1]) * ts[1] + (-Rt[a * 3 + 2]) * ts[2];
            for (int a = 0; a < 3; ++a)
                Pw[a] = Rt[a * 3 + 0] * Pc[0] + Rt[a * 3 + 1] * Pc[1] + Rt[a * 3 + 2] * Pc[2] + ti[a] * 1.0;
            float *pt = truep + ((size_t)y * cols + x) * 3;
            const float *pe = coords3d_est + ((size_t)y * cols + x) * 3;
            for (int c = 0; c < 3; ++c) {
                pt[c] = (float)Pw[c];
                const float s = pe[c] / pt[c];
                scales[(size_t)x * 3 * rows + 3 * (size_t)y + c] = s;
                if (fabsf(s) > 10) { scales[(size_t)x * 3 * rows + 3 * (size_t)y + c] = 0; outliers++; }
            }
        }
    double sum = 0;
    int inliers = 0;
    for (size_t i = 0; i < 3 * tot; ++i)
        if (scales[i] != 0 && scales[i] == scales[i]) { inliers++; sum += scales[i]; }
    const double scale = sum / (double)inliers;
    double sum_error = 0;
    inliers = 0;
    for (int x = 0; x < cols; ++x)
        for (int y = 0; y < rows; ++y) {
            const float *pe = coords3d_est + ((size_t)y * cols + x) * 3;
            const float *pt = truep + ((size_t)y * cols + x) * 3;
            const double e0 = pe[0] / scale, e1 = pe[1] / scale, e2 = pe[2] / scale;
            const double t0 = pt[0], t1 = pt[1], t2 = pt[2];
            const double d0 = e0 - t0, d1 = e1 - t1, d2 = e2 - t2;
            const double nrm = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
            if (e0 == e0 && e1 == e1 && e2 == e2 && t0 == t0 && t1 == t1 && t2 == t2)
                if (nrm < 50) { sum_error += nrm; inliers++; }
            if (error_image) {                                                     /* createErrorImage :583 */
                int q = 0;
                const double val = nrm * 255 / max_norm + 0.5;
                if (!trunc_to_int(val, &q)) q = (int)0x80000000;
                error_image[(size_t)y * cols + x] = (uint8_t)(q & 0xff);
            }
        }
    if (mean_scale) *mean_scale = scale;
    if (num_outliers) *num_outliers = outliers;
    if (points_used) *points_used = inliers;
    free(gt_depth); free(R); free(t); free(truep); free(scales);
    return sum_error * 1.0 / inliers;
}

/* ------------------------------------------------------------------------------------------ */
/* Camera::calculateTrueFlow                          camera.cc:209-249         */
/*                RsFrame::calculateImageCoordinatesRsFrame          rsframe.cc:740-768        */
/* ------------------------------------------------------------------------------------------ */
/* ux,uy,uz: unprojection maps of frame 1 (COLUMN-major rows x cols).  R2,t2: the (relative)
 * scanline poses of frame 2 (worldToCameraFrame's default argument, rsframe.h:239).  flow:
 * rows*cols*2 row-major (dx,dy).  O(rows^2 cols): every scanline pose is tried per pixel.
 * `best_row` is uninitialised in the reference when no displacement compares smaller than
 * INFINITY (all NaN); it starts at 0 here. */
ORC_API void orc_true_flow(const double *ux, const double *uy, const double *uz, const double *R2,
                           const double *t2, int rows, int cols, const double *K4, double *flow)
{
    const double fx = K4[0], cx = K4[2], cy = K4[3];
    for (int v = 0; v < rows; ++v)
        for (int u = 0; u < cols; ++u) {
            const size_t mi = (size_t)v + (size_t)u * rows;
            const double W[3] = {ux[mi], uy[mi], uz[mi]};
            double px = (double)u, py = (double)v;
            if (sqrt(W[0] * W[0] + W[1] * W[1] + W[2] * W[2]) != 0) {
                double min_diff = INFINITY, bx = 0, by = 0;
                int best = 0;
                for (int i = 0; i < rows; ++i) {
                    const double *R = R2 + 9 * i, *t = t2 + 3 * i;
                    const double Y = R[3] * W[0] + R[4] * W[1] + R[5] * W[2] + t[1] * 1.0;
                    const double Z = R[6] * W[0] + R[7] * W[1] + R[8] * W[2] + t[2] * 1.0;
                    const double qy = Y / Z * fx + cy;                       /* spaceToPlane: f_x for y too (Q12) */
                    const double diff = fabs(qy - (double)i);
                    if (diff < min_diff) { min_diff = diff; best = i; }
                }
                {
                    const double *R = R2 + 9 * best, *t = t2 + 3 * best;
                    const double X = R[0] * W[0] + R[1] * W[1] + R[2] * W[2] + t[0] * 1.0;
                    const double Y = R[3] * W[0] + R[4] * W[1] + R[5] * W[2] + t[1] * 1.0;
                    const double Z = R[6] * W[0] + R[7] * W[1] + R[8] * W[2] + t[2] * 1.0;
                    bx = X / Z * fx + cx;
                    by = Y / Z * fx + cy;
                }
                if (sqrt(bx * bx + by * by) != 0) { px = bx; py = by; }       /* camera.cc:236-238 */
            }
            flow[2 * ((size_t)v * cols + u)] = px - (double)u;
            flow[2 * ((size_t)v * cols + u) + 1] = py - (double)v;
        }
}
