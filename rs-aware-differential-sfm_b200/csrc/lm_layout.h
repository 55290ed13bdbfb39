// lm_layout.h -- HBM layout of the LM solver's per-residual-block inputs, shared by the kernels that
// fill it (refine.cu: k_refine_gather / k_depth_gather, preproc.cu: k_compact_scatter) and the solver.
//
// Tile-blocked structure of arrays, tile = kTile residual blocks:
//   blk[tile] = { xy[kTile] (x, y) | uu[kTile] (ux, uy) | aa[kTile] (alpha, alpha_k) } as double2,
// 12 KB contiguous, so that a whole tile arrives with ONE TMA bulk copy; the inverse depths live in
// separate planes d[2][tiles * kTile] (current point / candidate), 2 KB per tile.
#pragma once

#include <stddef.h>

namespace rsdsfm {

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
