// lm_layout.h -- HBM layout of the LM solver's per-residual-block inputs, shared by the kernels that
// fill it (refine.cu: k_refine_gather / k_depth_gather, preproc.cu: k_compact_scatter) and the solver.
//
// Tile-blocked structure of arrays, tile = kTile residual blocks:
//   blk[tile] = { xy[kTile] (x, y) | uu[kTile] (ux, uy) | aa[kTile] (alpha, alpha_k) } as double2,
// 12 KB contiguous, so that a whole tile arrives with ONE TMA bulk copy; the inverse depths live in
// separate planes d[2][tiles * kTile] (current point / candidate), 2 KB per tile.
//
// The model every one of those kernels evaluates -- RsResidual::operator() (nonlinearRefinement.cc:32-52), which the
// reference hands to Ceres as AutoDiffCostFunction<RsResidual,2,3,3,1,1> (nonlinearRefinement.cc:148-151, :215-216):
//     beta = 2/(2+k) * (alpha + k*alpha_k)
//     r    = u - beta * (A v d + B w)          A = [1 0 -x; 0 1 -y]
//                                              B = [-xy 1+x^2 -y; -(1+y^2) xy x]
// and its analytic Jacobian (pixel_fast / pixel_exact in lm_kernel.cuh):
//     dr/dv = -beta d A,  dr/dw = -beta B,  dr/dd = -beta A v =: e,  dr/dk = -(dbeta/dk)(A v d + B w),
//     dbeta/dk = 2/(2+k) * (alpha_k - (alpha + k alpha_k)/(2+k)).
#pragma once

#include <stddef.h>

namespace rsdsfm {

// the motion parameter block shared by the solver, the controller and the drivers
struct Motion {
    double v[3], w[3], k;
};

#ifndef RS_THREADS
#define RS_THREADS 256
#endif
constexpr int kTile = RS_THREADS;               // residual blocks per tile (one per thread of the solver CTA)

#ifdef __CUDACC__
__host__ __device__
#endif
inline size_t blk_index(int i, int field)
{
    return (size_t)(i / kTile) * (3 * kTile) + (size_t)field * kTile + (size_t)(i % kTile);
}

}  // namespace rsdsfm
