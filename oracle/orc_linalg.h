/*
 * orc_linalg.h -- small dense linear algebra for the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed by the
 * product path (rs-aware-differential-sfm_b200/); it is the checker the CUDA path is compared with.
 *
 * The reference delegates these to Eigen 3.3.4 (not vendored, not installed here):
 *   JacobiSVD<9x9>            minimal.cc:98-99
 *   EigenSolver<6x6>          minimal.cc:71-73
 *   SelfAdjointEigenSolver<3> minimal.cc:111-113
 *   MatrixXd::inverse()       minimal.cc:59,72
 *   LLT (inside Ceres' DenseSchurComplementSolver)
 * They are restated here from the published algorithms (one-sided Jacobi SVD, Householder
 * Hessenberg + Francis double-shift QR, cyclic Jacobi, LU with partial pivoting, Cholesky).
 * Row-major fixed-capacity storage, plain C99, no FMA contraction (build with -ffp-contract=off).
 */
#ifndef ORC_LINALG_H
#define ORC_LINALG_H

#include <math.h>
#include <string.h>

#define ORC_MAXN 9

/* ---- LU inverse with partial pivoting (Eigen: PartialPivLU based inverse()) ---- */
/* a: n x n row-major (stride n). Returns 0 on success, 1 if a zero pivot was met (result then
 * holds inf/nan exactly like a division by zero would give). */
static int orc_inverse(const double *a, int n, double *inv)
{
    double lu[ORC_MAXN * ORC_MAXN];
    int perm[ORC_MAXN];
    int singular = 0;
    memcpy(lu, a, sizeof(double) * n * n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int c = 0; c < n; ++c) {
        int p = c;
        double best = fabs(lu[c * n + c]);
        for (int r = c + 1; r < n; ++r) {
            double v = fabs(lu[r * n + c]);
            if (v > best) { best = v; p = r; }
        }
        if (p != c) {
            for (int j = 0; j < n; ++j) {
                double t = lu[c * n + j]; lu[c * n + j] = lu[p * n + j]; lu[p * n + j] = t;
            }
            int t = perm[c]; perm[c] = perm[p]; perm[p] = t;
        }
        if (lu[c * n + c] == 0.0) singular = 1;
        for (int r = c + 1; r < n; ++r) {
            lu[r * n + c] /= lu[c * n + c];
            double f = lu[r * n + c];
            for (int j = c + 1; j < n; ++j) lu[r * n + j] -= f * lu[c * n + j];
        }
    }
    /* solve for each unit vector */
    for (int col = 0; col < n; ++col) {
        double y[ORC_MAXN];
        for (int i = 0; i < n; ++i) {
            double s = (perm[i] == col) ? 1.0 : 0.0;
            for (int j = 0; j < i; ++j) s -= lu[i * n + j] * y[j];
            y[i] = s;
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = y[i];
            for (int j = i + 1; j < n; ++j) s -= lu[i * n + j] * inv[j * n + col];
            inv[i * n + col] = s / lu[i * n + i];
        }
    }
    return singular;
}

/* ---- one-sided (Hestenes) Jacobi SVD: right singular vectors of an m x n matrix, m >= n ---- */
/* a: m x n row-major, overwritten by U*diag(sigma).  v: n x n row-major right singular vectors.
 * sigma: n singular values (unsorted).  Returns number of sweeps used. */
static int orc_jacobi_svd(double *a, int m, int n, double *v, double *sigma)
{
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) v[i * n + j] = (i == j) ? 1.0 : 0.0;
    int sweep;
    for (sweep = 0; sweep < 60; ++sweep) {
        int rotated = 0;
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                double app = 0.0, aqq = 0.0, apq = 0.0;
                for (int i = 0; i < m; ++i) {
                    double x = a[i * n + p], y = a[i * n + q];
                    app += x * x; aqq += y * y; apq += x * y;
                }
                if (apq == 0.0) continue;
                if (fabs(apq) <= 1e-300 || fabs(apq) <= 2.220446049250313e-16 * sqrt(app) * sqrt(aqq) * 0.25)
                    continue;
                rotated = 1;
                double zeta = (aqq - app) / (2.0 * apq);
                double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t);
                double s = c * t;
                for (int i = 0; i < m; ++i) {
                    double x = a[i * n + p], y = a[i * n + q];
                    a[i * n + p] = c * x - s * y;
                    a[i * n + q] = s * x + c * y;
                }
                for (int i = 0; i < n; ++i) {
                    double x = v[i * n + p], y = v[i * n + q];
                    v[i * n + p] = c * x - s * y;
                    v[i * n + q] = s * x + c * y;
                }
            }
        }
        if (!rotated) break;
    }
    for (int j = 0; j < n; ++j) {
        double s = 0.0;
        for (int i = 0; i < m; ++i) s += a[i * n + j] * a[i * n + j];
        sigma[j] = sqrt(s);
    }
    return sweep;
}

/* ---- symmetric eigen-decomposition by cyclic Jacobi; eigenvalues ascending (Eigen convention) ---- */
/* a: n x n symmetric row-major (destroyed). evec: columns are unit eigenvectors. */
static void orc_sym_eig(double *a, int n, double *eval, double *evec)
{
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) evec[i * n + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, dia = 0.0;
        for (int i = 0; i < n; ++i) {
            dia += a[i * n + i] * a[i * n + i];
            for (int j = i + 1; j < n; ++j) off += a[i * n + j] * a[i * n + j];
        }
        if (off == 0.0 || off <= 1e-33 * dia) break;
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                double apq = a[p * n + q];
                if (apq == 0.0) continue;
                double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * apq);
                double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(1.0 + theta * theta));
                double c = 1.0 / sqrt(1.0 + t * t);
                double s = c * t;
                for (int k = 0; k < n; ++k) { /* A <- A*G */
                    double x = a[k * n + p], y = a[k * n + q];
                    a[k * n + p] = c * x - s * y;
                    a[k * n + q] = s * x + c * y;
                }
                for (int k = 0; k < n; ++k) { /* A <- G^T*A */
                    double x = a[p * n + k], y = a[q * n + k];
                    a[p * n + k] = c * x - s * y;
                    a[q * n + k] = s * x + c * y;
                }
                for (int k = 0; k < n; ++k) {
                    double x = evec[k * n + p], y = evec[k * n + q];
                    evec[k * n + p] = c * x - s * y;
                    evec[k * n + q] = s * x + c * y;
                }
            }
        }
    }
    for (int i = 0; i < n; ++i) eval[i] = a[i * n + i];
    /* selection sort ascending, permuting eigenvector columns */
    for (int i = 0; i < n - 1; ++i) {
        int k = i;
        for (int j = i + 1; j < n; ++j) if (eval[j] < eval[k]) k = j;
        if (k != i) {
            double t = eval[i]; eval[i] = eval[k]; eval[k] = t;
            for (int r = 0; r < n; ++r) {
                double u = evec[r * n + i]; evec[r * n + i] = evec[r * n + k]; evec[r * n + k] = u;
            }
        }
    }
}

/* ---- eigenvalues of a real general matrix: Householder Hessenberg + Francis double-shift QR ---- */
static double orc_sign(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

static void orc_hessenberg(double *a, int n)
{
    /* Householder similarity reduction to upper Hessenberg form, in place (row-major n x n). */
    for (int k = 0; k < n - 2; ++k) {
        double alpha = 0.0;
        for (int i = k + 1; i < n; ++i) alpha += a[i * n + k] * a[i * n + k];
        alpha = sqrt(alpha);
        if (alpha == 0.0) continue;
        double vv[ORC_MAXN];
        for (int i = 0; i < n; ++i) vv[i] = 0.0;
        double x0 = a[(k + 1) * n + k];
        double beta = (x0 >= 0.0) ? -alpha : alpha;
        vv[k + 1] = x0 - beta;
        for (int i = k + 2; i < n; ++i) vv[i] = a[i * n + k];
        double vnorm2 = 0.0;
        for (int i = k + 1; i < n; ++i) vnorm2 += vv[i] * vv[i];
        if (vnorm2 == 0.0) continue;
        /* A <- (I - 2 v v^T / v^T v) A */
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int i = k + 1; i < n; ++i) s += vv[i] * a[i * n + j];
            s = 2.0 * s / vnorm2;
            for (int i = k + 1; i < n; ++i) a[i * n + j] -= s * vv[i];
        }
        /* A <- A (I - 2 v v^T / v^T v) */
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            for (int j = k + 1; j < n; ++j) s += a[i * n + j] * vv[j];
            s = 2.0 * s / vnorm2;
            for (int j = k + 1; j < n; ++j) a[i * n + j] -= s * vv[j];
        }
        for (int i = k + 2; i < n; ++i) a[i * n + k] = 0.0;
    }
}

/* Returns 0 on success, 1 if the QR iteration did not converge (eigenvalues then NaN). */
static int orc_eigvals_general(const double *a_in, int n, double *wr, double *wi)
{
    double a[ORC_MAXN * ORC_MAXN];
    memcpy(a, a_in, sizeof(double) * n * n);
    for (int i = 0; i < n * n; ++i) {
        if (!isfinite(a[i])) {
            for (int j = 0; j < n; ++j) { wr[j] = NAN; wi[j] = NAN; }
            return 1;
        }
    }
    orc_hessenberg(a, n);
#define A_(i, j) a[(i) * n + (j)]
    double anorm = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = (i > 0 ? i - 1 : 0); j < n; ++j) anorm += fabs(A_(i, j));
    int nn = n - 1;
    double t = 0.0;
    double p = 0, q = 0, r = 0, s, w, x, y, z;
    while (nn >= 0) {
        int its = 0, l;
        do {
            for (l = nn; l >= 1; --l) {
                s = fabs(A_(l - 1, l - 1)) + fabs(A_(l, l));
                if (s == 0.0) s = anorm;
                if (fabs(A_(l, l - 1)) + s == s) { A_(l, l - 1) = 0.0; break; }
            }
            x = A_(nn, nn);
            if (l == nn) {                      /* one real root */
                wr[nn] = x + t; wi[nn] = 0.0; nn -= 1;
            } else {
                y = A_(nn - 1, nn - 1);
                w = A_(nn, nn - 1) * A_(nn - 1, nn);
                if (l == nn - 1) {              /* a 2x2 block: two roots */
                    p = 0.5 * (y - x);
                    q = p * p + w;
                    z = sqrt(fabs(q));
                    x += t;
                    if (q >= 0.0) {
                        z = p + orc_sign(z, p);
                        wr[nn - 1] = wr[nn] = x + z;
                        if (z != 0.0) wr[nn] = x - w / z;
                        wi[nn - 1] = wi[nn] = 0.0;
                    } else {
                        wr[nn - 1] = wr[nn] = x + p;
                        wi[nn - 1] = z; wi[nn] = -z;
                    }
                    nn -= 2;
                } else {                        /* no root yet: one Francis step */
                    if (its == 60) {
                        for (int j = 0; j < n; ++j) { wr[j] = NAN; wi[j] = NAN; }
                        return 1;
                    }
                    if (its == 10 || its == 20) {   /* exceptional shift */
                        t += x;
                        for (int i = 0; i <= nn; ++i) A_(i, i) -= x;
                        s = fabs(A_(nn, nn - 1)) + fabs(A_(nn - 1, nn - 2));
                        y = x = 0.75 * s;
                        w = -0.4375 * s * s;
                    }
                    ++its;
                    int m;
                    for (m = nn - 2; m >= l; --m) {
                        z = A_(m, m);
                        r = x - z; s = y - z;
                        p = (r * s - w) / A_(m + 1, m) + A_(m, m + 1);
                        q = A_(m + 1, m + 1) - z - r - s;
                        r = A_(m + 2, m + 1);
                        s = fabs(p) + fabs(q) + fabs(r);
                        p /= s; q /= s; r /= s;
                        if (m == l) break;
                        double u = fabs(A_(m, m - 1)) * (fabs(q) + fabs(r));
                        double v = fabs(p) * (fabs(A_(m - 1, m - 1)) + fabs(z) + fabs(A_(m + 1, m + 1)));
                        if (u + v == v) break;
                    }
                    for (int i = m + 2; i <= nn; ++i) {
                        A_(i, i - 2) = 0.0;
                        if (i != m + 2) A_(i, i - 3) = 0.0;
                    }
                    for (int k = m; k <= nn - 1; ++k) {
                        if (k != m) {
                            p = A_(k, k - 1);
                            q = A_(k + 1, k - 1);
                            r = 0.0;
                            if (k != nn - 1) r = A_(k + 2, k - 1);
                            x = fabs(p) + fabs(q) + fabs(r);
                            if (x != 0.0) { p /= x; q /= x; r /= x; }
                        }
                        s = orc_sign(sqrt(p * p + q * q + r * r), p);
                        if (s != 0.0) {
                            if (k == m) {
                                if (l != m) A_(k, k - 1) = -A_(k, k - 1);
                            } else {
                                A_(k, k - 1) = -s * x;
                            }
                            p += s;
                            x = p / s; y = q / s; z = r / s;
                            q /= p; r /= p;
                            for (int j = k; j <= nn; ++j) {
                                p = A_(k, j) + q * A_(k + 1, j);
                                if (k != nn - 1) { p += r * A_(k + 2, j); A_(k + 2, j) -= p * z; }
                                A_(k + 1, j) -= p * y;
                                A_(k, j) -= p * x;
                            }
                            int mmin = nn < k + 3 ? nn : k + 3;
                            for (int i = l; i <= mmin; ++i) {
                                p = x * A_(i, k) + y * A_(i, k + 1);
                                if (k != nn - 1) { p += z * A_(i, k + 2); A_(i, k + 2) -= p * r; }
                                A_(i, k + 1) -= p * q;
                                A_(i, k) -= p;
                            }
                        }
                    }
                }
            }
        } while (l < nn - 1);
    }
#undef A_
    return 0;
}

/* ---- Cholesky (LL^T) solve of an n x n SPD system, n <= 9.  Returns 0 ok, 1 not PD ---- */
static int orc_cholesky_solve(const double *a, int n, const double *b, double *x)
{
    double l[ORC_MAXN * ORC_MAXN];
    memset(l, 0, sizeof l);
    for (int j = 0; j < n; ++j) {
        double d = a[j * n + j];
        for (int k = 0; k < j; ++k) d -= l[j * n + k] * l[j * n + k];
        if (!(d > 0.0)) return 1;           /* also catches NaN */
        d = sqrt(d);
        l[j * n + j] = d;
        for (int i = j + 1; i < n; ++i) {
            double s = a[i * n + j];
            for (int k = 0; k < j; ++k) s -= l[i * n + k] * l[j * n + k];
            l[i * n + j] = s / d;
        }
    }
    double y[ORC_MAXN];
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
    return 0;
}

#endif /* ORC_LINALG_H */
