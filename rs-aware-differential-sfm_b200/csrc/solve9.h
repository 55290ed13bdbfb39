// solve9.h -- the 9-point RS differential epipolar solver (minimal::calculateVelocities,
// minimal.cc:36-177) with the small dense factorisations it needs.
//
// The reference delegates the factorisations to Eigen 3.3.4 (JacobiSVD<9x9> :98, EigenSolver<6x6>
// :71-73, SelfAdjointEigenSolver<3x3> :111, MatrixXd::inverse() :59,72), which is not vendored;
// they are implemented here from the textbook algorithms as fixed-size __host__ __device__
// templates so a hypothesis can be fitted by one thread (host call today, one thread per
// hypothesis on the device for the sequence pipeline).
#pragma once

#include <math.h>

#ifdef __CUDACC__
#define S9_HD __host__ __device__
#else
#define S9_HD
#endif

namespace rsdsfm {
namespace s9 {

template <int N>
struct Mat {                       // row-major N x N
    double a[N * N];
    S9_HD double &operator()(int r, int c) { return a[r * N + c]; }
    S9_HD double operator()(int r, int c) const { return a[r * N + c]; }
};

template <int N>
S9_HD inline Mat<N> mul(const Mat<N> &x, const Mat<N> &y)
{
    Mat<N> z;
    for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) {
            double s = 0.0;
            for (int k = 0; k < N; ++k) s += x(r, k) * y(k, c);
            z(r, c) = s;
        }
    return z;
}
template <int N>
S9_HD inline Mat<N> transpose(const Mat<N> &x)
{
    Mat<N> z;
    for (int r = 0; r < N; ++r) for (int c = 0; c < N; ++c) z(r, c) = x(c, r);
    return z;
}

// Inverse by LU with partial pivoting.
template <int N>
S9_HD inline Mat<N> inverse(const Mat<N> &m)
{
    Mat<N> lu = m, inv;
    int perm[N];
    for (int i = 0; i < N; ++i) perm[i] = i;
    for (int c = 0; c < N; ++c) {
        int piv = c;
        double best = fabs(lu(c, c));
        for (int r = c + 1; r < N; ++r) if (fabs(lu(r, c)) > best) { best = fabs(lu(r, c)); piv = r; }
        if (piv != c) {
            for (int j = 0; j < N; ++j) { double t = lu(c, j); lu(c, j) = lu(piv, j); lu(piv, j) = t; }
            int t = perm[c]; perm[c] = perm[piv]; perm[piv] = t;
        }
        for (int r = c + 1; r < N; ++r) {
            lu(r, c) /= lu(c, c);
            const double f = lu(r, c);
            for (int j = c + 1; j < N; ++j) lu(r, j) -= f * lu(c, j);
        }
    }
    for (int col = 0; col < N; ++col) {
        double y[N];
        for (int i = 0; i < N; ++i) {
            double s = (perm[i] == col) ? 1.0 : 0.0;
            for (int j = 0; j < i; ++j) s -= lu(i, j) * y[j];
            y[i] = s;
        }
        for (int i = N - 1; i >= 0; --i) {
            double s = y[i];
            for (int j = i + 1; j < N; ++j) s -= lu(i, j) * inv(j, col);
            inv(i, col) = s / lu(i, i);
        }
    }
    return inv;
}

// Right singular vector of the smallest singular value (one-sided Jacobi on the columns).
template <int N>
S9_HD inline void smallest_right_singular_vector(Mat<N> a, double *out)
{
    Mat<N> v;
    for (int r = 0; r < N; ++r) for (int c = 0; c < N; ++c) v(r, c) = (r == c) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                double app = 0, aqq = 0, apq = 0;
                for (int i = 0; i < N; ++i) { app += a(i, p) * a(i, p); aqq += a(i, q) * a(i, q); apq += a(i, p) * a(i, q); }
                if (apq == 0.0 || fabs(apq) <= 1e-300 || fabs(apq) <= 5.551115123125783e-17 * sqrt(app) * sqrt(aqq)) continue;
                rotated = true;
                const double zeta = (aqq - app) / (2.0 * apq);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < N; ++i) {
                    const double x = a(i, p), y = a(i, q);
                    a(i, p) = c * x - s * y; a(i, q) = s * x + c * y;
                    const double vx = v(i, p), vy = v(i, q);
                    v(i, p) = c * vx - s * vy; v(i, q) = s * vx + c * vy;
                }
            }
        if (!rotated) break;
    }
    int best = 0;
    double bestn = INFINITY;
    for (int c = 0; c < N; ++c) {
        double s = 0;
        for (int i = 0; i < N; ++i) s += a(i, c) * a(i, c);
        if (s < bestn) { bestn = s; best = c; }
    }
    for (int i = 0; i < N; ++i) out[i] = v(i, best);
}

// Symmetric eigen-decomposition (cyclic Jacobi), eigenvalues ascending, unit eigenvector columns.
template <int N>
S9_HD inline void sym_eig(Mat<N> a, double *eval, Mat<N> &evec)
{
    for (int r = 0; r < N; ++r) for (int c = 0; c < N; ++c) evec(r, c) = (r == c) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0, dia = 0;
        for (int i = 0; i < N; ++i) { dia += a(i, i) * a(i, i); for (int j = i + 1; j < N; ++j) off += a(i, j) * a(i, j); }
        if (off == 0.0 || off <= 1e-33 * dia) break;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                const double apq = a(p, q);
                if (apq == 0.0) continue;
                const double th = (a(q, q) - a(p, p)) / (2.0 * apq);
                const double t = (th >= 0.0 ? 1.0 : -1.0) / (fabs(th) + sqrt(1.0 + th * th));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int k = 0; k < N; ++k) { const double x = a(k, p), y = a(k, q); a(k, p) = c * x - s * y; a(k, q) = s * x + c * y; }
                for (int k = 0; k < N; ++k) { const double x = a(p, k), y = a(q, k); a(p, k) = c * x - s * y; a(q, k) = s * x + c * y; }
                for (int k = 0; k < N; ++k) { const double x = evec(k, p), y = evec(k, q); evec(k, p) = c * x - s * y; evec(k, q) = s * x + c * y; }
            }
    }
    for (int i = 0; i < N; ++i) eval[i] = a(i, i);
    for (int i = 0; i < N - 1; ++i) {
        int k = i;
        for (int j = i + 1; j < N; ++j) if (eval[j] < eval[k]) k = j;
        if (k != i) {
            const double t = eval[i]; eval[i] = eval[k]; eval[k] = t;
            for (int r = 0; r < N; ++r) { const double u = evec(r, i); evec(r, i) = evec(r, k); evec(r, k) = u; }
        }
    }
}

S9_HD inline double sgn(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

// Eigenvalues of a real general matrix: Householder reduction to Hessenberg form followed by the
// Francis double-shift QR iteration.  Returns false when the iteration does not converge.
template <int N>
S9_HD inline bool eigvals_general(Mat<N> a, double *wr, double *wi)
{
    for (int i = 0; i < N * N; ++i)
        if (!isfinite(a.a[i])) { for (int j = 0; j < N; ++j) { wr[j] = NAN; wi[j] = NAN; } return false; }
    for (int k = 0; k < N - 2; ++k) {          // Hessenberg
        double alpha = 0;
        for (int i = k + 1; i < N; ++i) alpha += a(i, k) * a(i, k);
        alpha = sqrt(alpha);
        if (alpha == 0.0) continue;
        double v[N];
        for (int i = 0; i < N; ++i) v[i] = 0.0;
        const double x0 = a(k + 1, k);
        v[k + 1] = x0 - ((x0 >= 0.0) ? -alpha : alpha);
        for (int i = k + 2; i < N; ++i) v[i] = a(i, k);
        double vn = 0;
        for (int i = k + 1; i < N; ++i) vn += v[i] * v[i];
        if (vn == 0.0) continue;
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int i = k + 1; i < N; ++i) s += v[i] * a(i, j);
            s = 2.0 * s / vn;
            for (int i = k + 1; i < N; ++i) a(i, j) -= s * v[i];
        }
        for (int i = 0; i < N; ++i) {
            double s = 0;
            for (int j = k + 1; j < N; ++j) s += a(i, j) * v[j];
            s = 2.0 * s / vn;
            for (int j = k + 1; j < N; ++j) a(i, j) -= s * v[j];
        }
        for (int i = k + 2; i < N; ++i) a(i, k) = 0.0;
    }
    double anorm = 0;
    for (int i = 0; i < N; ++i) for (int j = (i > 0 ? i - 1 : 0); j < N; ++j) anorm += fabs(a(i, j));
    int nn = N - 1;
    double t = 0, p = 0, q = 0, r = 0, s, w, x, y, z;
    while (nn >= 0) {
        int its = 0, l;
        do {
            for (l = nn; l >= 1; --l) {
                s = fabs(a(l - 1, l - 1)) + fabs(a(l, l));
                if (s == 0.0) s = anorm;
                if (fabs(a(l, l - 1)) + s == s) { a(l, l - 1) = 0.0; break; }
            }
            x = a(nn, nn);
            if (l == nn) { wr[nn] = x + t; wi[nn] = 0.0; nn -= 1; }
            else {
                y = a(nn - 1, nn - 1);
                w = a(nn, nn - 1) * a(nn - 1, nn);
                if (l == nn - 1) {
                    p = 0.5 * (y - x); q = p * p + w; z = sqrt(fabs(q)); x += t;
                    if (q >= 0.0) {
                        z = p + sgn(z, p);
                        wr[nn - 1] = wr[nn] = x + z;
                        if (z != 0.0) wr[nn] = x - w / z;
                        wi[nn - 1] = wi[nn] = 0.0;
                    } else { wr[nn - 1] = wr[nn] = x + p; wi[nn - 1] = z; wi[nn] = -z; }
                    nn -= 2;
                } else {
                    if (its == 60) { for (int j = 0; j < N; ++j) { wr[j] = NAN; wi[j] = NAN; } return false; }
                    if (its == 10 || its == 20) {
                        t += x;
                        for (int i = 0; i <= nn; ++i) a(i, i) -= x;
                        s = fabs(a(nn, nn - 1)) + fabs(a(nn - 1, nn - 2));
                        y = x = 0.75 * s; w = -0.4375 * s * s;
                    }
                    ++its;
                    int m;
                    for (m = nn - 2; m >= l; --m) {
                        z = a(m, m); r = x - z; s = y - z;
                        p = (r * s - w) / a(m + 1, m) + a(m, m + 1);
                        q = a(m + 1, m + 1) - z - r - s;
                        r = a(m + 2, m + 1);
                        s = fabs(p) + fabs(q) + fabs(r);
                        p /= s; q /= s; r /= s;
                        if (m == l) break;
                        const double u = fabs(a(m, m - 1)) * (fabs(q) + fabs(r));
                        const double vv = fabs(p) * (fabs(a(m - 1, m - 1)) + fabs(z) + fabs(a(m + 1, m + 1)));
                        if (u + vv == vv) break;
                    }
                    for (int i = m + 2; i <= nn; ++i) { a(i, i - 2) = 0.0; if (i != m + 2) a(i, i - 3) = 0.0; }
                    for (int k = m; k <= nn - 1; ++k) {
                        if (k != m) {
                            p = a(k, k - 1); q = a(k + 1, k - 1); r = 0.0;
                            if (k != nn - 1) r = a(k + 2, k - 1);
                            x = fabs(p) + fabs(q) + fabs(r);
                            if (x != 0.0) { p /= x; q /= x; r /= x; }
                        }
                        s = sgn(sqrt(p * p + q * q + r * r), p);
                        if (s != 0.0) {
                            if (k == m) { if (l != m) a(k, k - 1) = -a(k, k - 1); }
                            else a(k, k - 1) = -s * x;
                            p += s; x = p / s; y = q / s; z = r / s; q /= p; r /= p;
                            for (int j = k; j <= nn; ++j) {
                                p = a(k, j) + q * a(k + 1, j);
                                if (k != nn - 1) { p += r * a(k + 2, j); a(k + 2, j) -= p * z; }
                                a(k + 1, j) -= p * y; a(k, j) -= p * x;
                            }
                            const int mmin = nn < k + 3 ? nn : k + 3;
                            for (int i = l; i <= mmin; ++i) {
                                p = x * a(i, k) + y * a(i, k + 1);
                                if (k != nn - 1) { p += z * a(i, k + 2); a(i, k + 2) -= p * r; }
                                a(i, k + 1) -= p * q; a(i, k) -= p;
                            }
                        }
                    }
                }
            }
        } while (l < nn - 1);
    }
    return true;
}

S9_HD inline Mat<3> rot_y(double ang)
{
    const double c = cos(ang), s = sin(ang);
    Mat<3> r; r(0,0)=c; r(0,1)=0; r(0,2)=s; r(1,0)=0; r(1,1)=1; r(1,2)=0; r(2,0)=-s; r(2,1)=0; r(2,2)=c;
    return r;
}
S9_HD inline Mat<3> rot_z(double ang)
{
    const double c = cos(ang), s = sin(ang);
    Mat<3> r; r(0,0)=c; r(0,1)=-s; r(0,2)=0; r(1,0)=s; r(1,1)=c; r(1,2)=0; r(2,0)=0; r(2,1)=0; r(2,2)=1;
    return r;
}

// minimal::calculateVelocities.  q,u: 2x9 interleaved; out7 = (w, v, k).
S9_HD inline void calculate_velocities(const double *q, const double *u, const double *alpha, const double *alpha_k,
                                       bool use_alpha_k, double *out7)
{
    const double kThresholdLambda = 0.000001, kTolImag = 0.00001, kPi = 3.14159265358979323846;
    double k = 0.0, beta[9];
    Mat<9> z;
    for (int i = 0; i < 9; ++i) {
        const double qx = q[2 * i], qy = q[2 * i + 1], ux = u[2 * i], uy = u[2 * i + 1];
        z(i, 0) = -uy; z(i, 1) = ux; z(i, 2) = uy * qx - ux * qy;
        z(i, 3) = qx * qx; z(i, 4) = 2.0 * qx * qy; z(i, 5) = 2.0 * qx;
        z(i, 6) = qy * qy; z(i, 7) = 2 * qy; z(i, 8) = 1.0;
        beta[i] = alpha[i];
    }
    if (use_alpha_k) {
        // k from det Z(k) = 0, reduced to the eigenvalues of P * P_k^-1  (:58-80)
        Mat<3> a;
        Mat<6> efhj;
        double dg[6][3], bc[3][6];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a(i, j) = z(i, j);
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) efhj(i, j) = z(3 + i, 3 + j);
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 3; ++j) dg[i][j] = z(3 + i, j);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 6; ++j) bc[i][j] = z(i, 3 + j);
        const Mat<3> ai = inverse(a);
        double dga[6][3];
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 3; ++j) dga[i][j] = dg[i][0] * ai(0, j) + dg[i][1] * ai(1, j) + dg[i][2] * ai(2, j);
        Mat<6> p, pk;
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) {
                double s = 0, sk = 0;
                for (int c = 0; c < 3; ++c) { s += dga[i][c] * alpha[c] * bc[c][j]; sk += dga[i][c] * alpha_k[c] * bc[c][j]; }
                p(i, j) = alpha[3 + i] * efhj(i, j) - s;
                pk(i, j) = alpha_k[3 + i] * efhj(i, j) - sk;
            }
        const Mat<6> m = mul(p, inverse(pk));
        double wr[6], wi[6];
        eigvals_general(m, wr, wi);
        k = INFINITY;
        for (int i = 0; i < 6; ++i)
            if (fabs(wi[i]) < kTolImag && fabs(wr[i]) < fabs(k)) k = wr[i];
        for (int i = 0; i < 9; ++i) beta[i] = (alpha[i] + k * alpha_k[i]) * (2.0 / (2.0 + k));
    }
    for (int i = 0; i < 9; ++i) for (int c = 3; c < 9; ++c) z(i, c) *= beta[i];

    double e[9];
    bool finite = true;
    for (int i = 0; i < 81; ++i) if (!isfinite(z.a[i])) finite = false;
    if (finite) smallest_right_singular_vector(z, e);
    else for (int i = 0; i < 9; ++i) e[i] = NAN;
    const double norm_v0 = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    for (int i = 0; i < 9; ++i) e[i] = e[i] / norm_v0;
    const double v0[3] = {e[0], e[1], e[2]};
    Mat<3> s;
    s(0,0)=e[3]; s(0,1)=e[4]; s(0,2)=e[5]; s(1,0)=e[4]; s(1,1)=e[6]; s(1,2)=e[7]; s(2,0)=e[5]; s(2,1)=e[7]; s(2,2)=e[8];
    double lamb[3];
    Mat<3> v1;
    finite = true;
    for (int i = 0; i < 9; ++i) if (!isfinite(s.a[i])) finite = false;
    if (finite) sym_eig(s, lamb, v1);
    else { for (int i = 0; i < 3; ++i) lamb[i] = NAN; for (int i = 0; i < 9; ++i) v1.a[i] = NAN; }
    for (int r = 0; r < 3; ++r) { const double t = v1(r, 0); v1(r, 0) = v1(r, 2); v1(r, 2) = t; }   // :114
    const double sigma0 = (2 * lamb[2] + lamb[1] - lamb[0]) / 3;
    const double sigma1 = (lamb[2] + 2 * lamb[1] + lamb[0]) / 3;
    const double sigma2 = (-lamb[2] + lamb[1] + 2 * lamb[0]) / 3;
    const double lambda = sigma0 - sigma2;
    double theta = 0;
    if (!(lambda < kThresholdLambda)) theta = acos(-sigma1 / lambda);
    const Mat<3> r_v = rot_y((theta - kPi) / 2), r_u = rot_y(theta);
    const Mat<3> v_ = mul(v1, transpose(r_v));
    Mat<3> nv = v_;
    for (int i = 0; i < 9; ++i) nv.a[i] = -nv.a[i];
    const Mat<3> u_ = mul(nv, r_u);
    Mat<3> sig1, sigl;
    for (int i = 0; i < 9; ++i) sig1.a[i] = 0.0;
    sig1(0, 0) = 1; sig1(1, 1) = 1;
    for (int i = 0; i < 9; ++i) sigl.a[i] = lambda * sig1.a[i];
    const Mat<3> rz1 = rot_z(kPi / 2), rz2 = rot_z(-kPi / 2);
    const Mat<3> *base[4] = {&v_, &v_, &u_, &u_};
    const Mat<3> *rz[4] = {&rz1, &rz2, &rz1, &rz2};
    int index_max = 0;
    double best = 0;
    for (int c = 0; c < 4; ++c) {
        const Mat<3> vh = mul(mul(mul(*base[c], *rz[c]), sig1), transpose(*base[c]));
        const double d = vh(2, 1) * v0[0] + vh(0, 2) * v0[1] + vh(1, 0) * v0[2];
        if (c == 0 || d > best) { best = d; index_max = c; }
    }
    // the omega paired with the optimal v_hat uses the OTHER basis (:158-171)
    const Mat<3> &ob = (index_max < 2) ? u_ : v_;
    const Mat<3> &orz = (index_max % 2 == 0) ? rz1 : rz2;
    const Mat<3> wh = mul(mul(mul(ob, orz), sigl), transpose(ob));
    out7[0] = wh(2, 1); out7[1] = wh(0, 2); out7[2] = wh(1, 0);
    out7[3] = v0[0]; out7[4] = v0[1]; out7[5] = v0[2];
    out7[6] = k;
}

}  // namespace s9
}  // namespace rsdsfm
