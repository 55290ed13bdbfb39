// rs_math.cuh -- the motion parameter block shared by the solver, the controller and the drivers,
// and the model they all evaluate.
//
// RsResidual::operator() (nonlinearRefinement.cc:32-52), evaluated by the reference through Ceres'
// AutoDiffCostFunction<RsResidual,2,3,3,1,1> (nonlinearRefinement.cc:148-151, :215-216):
//     beta = 2/(2+k) * (alpha + k*alpha_k)
//     r    = u - beta * (A v d + B w)          A = [1 0 -x; 0 1 -y]
//                                              B = [-xy 1+x^2 -y; -(1+y^2) xy x]
// Analytic Jacobian used by px_eval / ft_times in refine.cu:
//     dr/dv = -beta d A,  dr/dw = -beta B,  dr/dd = -beta A v =: e,  dr/dk = -(dbeta/dk)(A v d + B w),
//     dbeta/dk = 2/(2+k) * (alpha_k - (alpha + k alpha_k)/(2+k)).
#pragma once

#include "lm_controller.h"

namespace rsdsfm {

struct Motion {
    double v[3], w[3], k;
};

}  // namespace rsdsfm
