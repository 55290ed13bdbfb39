// rs_math.cuh -- per-correspondence RS-aware flow residual and its analytic Jacobian.
//
// Replaces RsResidual::operator() (nonlinearRefinement.cc:32-52) evaluated through Ceres'
// AutoDiffCostFunction<RsResidual,2,3,3,1,1> (nonlinearRefinement.cc:148-151, :215-216):
//     beta = 2/(2+k) * (alpha + k*alpha_k)
//     r    = u - beta * (A v d + B w)          A = [1 0 -x; 0 1 -y]
//                                              B = [-xy 1+x^2 -y; -(1+y^2) xy x]
// Jacobian: dr/dv = -beta d A, dr/dw = -beta B, dr/dd = -beta A v =: e, dr/dk = -(dbeta/dk)(A v d + B w).
#pragma once

#include "lm_controller.h"

namespace rsdsfm {

struct Motion {
    double v[3], w[3], k;
};

// Observation of one correspondence (normalised coords, gamma-scaled normalised flow, RS factors).
struct Obs {
    double x, y, ux, uy, alpha, alpha_k;
};

// Residual only.
__device__ __forceinline__ void rs_residual(const Obs &o, const Motion &m, double c2, double d, double &r0, double &r1)
{
    const double beta = c2 * (o.alpha + m.k * o.alpha_k);
    const double xy = o.x * o.y;
    const double ex0 = d * (m.v[0] - o.x * m.v[2]) + (-xy * m.w[0] + (1.0 + o.x * o.x) * m.w[1] - o.y * m.w[2]);
    const double ex1 = d * (m.v[1] - o.y * m.v[2]) + (-(1.0 + o.y * o.y) * m.w[0] + xy * m.w[1] + o.x * m.w[2]);
    r0 = o.ux - beta * ex0;
    r1 = o.uy - beta * ex1;
}

// Residual, depth column e (2) and the NF free motion columns (NF = 0: none, 6: v,w, 7: v,w,k).
template <int NF>
__device__ __forceinline__ void rs_residual_jac(const Obs &o, const Motion &m, double c2, double d, double &r0,
                                                double &r1, double &e0, double &e1, double (&F0)[NF > 0 ? NF : 1],
                                                double (&F1)[NF > 0 ? NF : 1])
{
    const double x = o.x, y = o.y;
    const double ak = o.alpha + m.k * o.alpha_k;
    const double beta = c2 * ak;
    const double a0 = m.v[0] - x * m.v[2];           // A v
    const double a1 = m.v[1] - y * m.v[2];
    const double xy = x * y, xx1 = 1.0 + x * x, yy1 = 1.0 + y * y;
    const double b0 = -xy * m.w[0] + xx1 * m.w[1] - y * m.w[2];   // B w
    const double b1 = -yy1 * m.w[0] + xy * m.w[1] + x * m.w[2];
    const double p0 = d * a0 + b0, p1 = d * a1 + b1;
    r0 = o.ux - beta * p0;
    r1 = o.uy - beta * p1;
    e0 = -beta * a0;
    e1 = -beta * a1;
    if (NF >= 6) {
        const double bd = beta * d;
        F0[0] = -bd;  F0[1] = 0.0;  F0[2] = bd * x;
        F1[0] = 0.0;  F1[1] = -bd;  F1[2] = bd * y;
        F0[3] = beta * xy;   F0[4] = -beta * xx1;  F0[5] = beta * y;
        F1[3] = beta * yy1;  F1[4] = -beta * xy;   F1[5] = -beta * x;
    }
    if (NF == 7) {
        // d beta / d k = c2 * (alpha_k - (alpha + k alpha_k) / (2 + k)),  1/(2+k) = c2/2
        const double dbeta = c2 * (o.alpha_k - ak * (0.5 * c2));
        F0[6] = -dbeta * p0;
        F1[6] = -dbeta * p1;
    }
}

}  // namespace rsdsfm
