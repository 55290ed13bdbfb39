// stages.h -- the device-pointer stage functions implemented by the translation units
// (preproc.cu, ransac.cu, refine.cu, rectify.cu) and composed by api.cu / pipeline.cu.
// Every function queues work on ctx->stream; those that return a count or a host-side result
// synchronise that stream, the others do not.
#pragma once

#include "common.cuh"
#include "lm_controller.h"
#include "lm_controller.h"
#include "lm_layout.h"

namespace rsdsfm {

// a2: flatten (main.cc:398-432)  [synchronises: *n_out]
int flatten_device(rsdsfm_ctx *, const double *flow_img, int rows, int cols, const double *K4, double gamma, double thr,
                   double *coord, double *flow, double *coord_px, double *flow_px, int32_t *pixel_index, int *n_out);
// a3/a4: getAlpha / getAlphaK (minimal.cc:179-197)
int alpha_device(rsdsfm_ctx *, const double *flow_px, const double *q_px, int n, double h, double gamma, double *alpha,
                 double *alpha_k);
// a5..a8: hypothesis fit on sampled 9-tuples and batch scoring (minimal.cc:209-306)  [synchronise]
int ransac_fit_device(rsdsfm_ctx *, const double *q, const double *u, const double *alpha, const double *alpha_k, int n,
                      int use_alpha_k, const int32_t *samples_host, int H, double *hyps7_out);
int ransac_score_device(rsdsfm_ctx *, const double *q, const double *u, const double *alpha, const double *alpha_k, int n,
                        const double *hyps7, int H, double tol, int *counts, double *sumerr, int *best_idx,
                        uint8_t *mask_best, double *inv_depth_best);
// a6 tail: consensus-set gather  [synchronises: *m_out]
int gather_inliers_device(rsdsfm_ctx *, const double *q, const double *alpha, const double *alpha_k, int n,
                          const uint8_t *mask, const double *inv_depth, double *inliers3, double *alpha_in,
                          double *alpha_k_in, int32_t *index_in, int *m_out);
// a8: estimateInverseDepths (nonlinearRefinement.cc:109-180)  [synchronises]
int estimate_inverse_depths_device(rsdsfm_ctx *, const double *coord, const double *flow, int n, const double *v,
                                   const double *w, double k, const double *alpha, const double *alpha_k,
                                   double *inv_depth, rsdsfm_lm_summary *summary);
// a9: nonLinearRefinement (nonlinearRefinement.cc:183-252)
int refine_device(rsdsfm_ctx *, const double *flow, const double *inliers3, const double *alpha, const double *alpha_k,
                  int m, double *v, double *w, double *k, int const_acc, const int32_t *flow_index,
                  const rsdsfm_lm_options *, double *z_out, rsdsfm_lm_summary *);                        // [synchronises]
int refine_async(rsdsfm_ctx *, const double *flow, const double *inliers3, const double *alpha, const double *alpha_k,
                 int m, const double *v, const double *w, double k, int const_acc, const int32_t *flow_index,
                 const rsdsfm_lm_options *, double *z_out, double *zstats_rows = nullptr);   // zstats_rows: num_sms x 3, see glue_device
// the same solve on inputs a producer kernel wrote straight into the solver's layout (compact step inputs)
int lm_input_buffers(rsdsfm_ctx *, int m, void **blk, double **d0, int **input_flag);
int refine_prepared_async(rsdsfm_ctx *, int m, const double *v, const double *w, double k, int const_acc,
                          const rsdsfm_lm_options *, const double *z_in, double *z_out, double *zstats_rows);
// flow field + RANSAC winner (mask, inverse depths over the flattened points) -> solver layout (preproc.cu)
int compact_build_device(rsdsfm_ctx *, const void *flow_img, int flow_f32, int rows, int cols, const double *K4, double gamma,
                         double thr, const uint8_t *mask, const double *inv_depth, int n, int m, void *blk, double *d0,
                         double *z_in, double *xy, int *input_flag);
size_t lm_mailbox_bytes();                              // mailbox of a row split over GPUs (lm_kernel.cuh)
int lm_grid_size(const rsdsfm_ctx *);                   // CTAs of the LM kernel = rows of its z statistics
int lm_reserve(rsdsfm_ctx *, int m);                     // pre-sizes the solver's buffers for up to m residual blocks
int lm_collect_enqueue(rsdsfm_ctx *, const double *stats_dev8);   // zero-copy read-back into the I/O slot's pinned area
int lm_collect_finish(rsdsfm_ctx *, int nf, int m, Motion *mot, rsdsfm_lm_summary *, bool *overflow);
int lm_collect(rsdsfm_ctx *, int nf, int m, Motion *mot, rsdsfm_lm_summary *, bool *overflow);            // [synchronises]
const double *lm_motion_device(rsdsfm_ctx *);
// a10/a11: sign fix + depth raster (main.cc:466-509)
int glue_device(rsdsfm_ctx *, double *z, int zs, const double *xyz, int xs, int m, const double *K4, int rows, int cols,
                double z_min_init, int layout, double *depth_map, uint8_t *depth_img, double *stats_dev8,
                const double *zrows = nullptr, int nzrows = 0);   // zrows: {sum z, max z, max -z} per producer CTA, if already known
// a12: setRelativePose (rsframe.cc:771-800), motion and depth statistics read on the device
int poses_device(rsdsfm_ctx *, const double *motion7_dev, const double *stats_dev, double gamma, int rows, double *R,
                 double *t);
// a13/a14: backProject(Gs) (rsframe.cc:803-878)
int backproject_device(rsdsfm_ctx *, const uint8_t *image, const double *depth, int layout, int rows, int cols,
                       const double *K4, const double *R, const double *t, int gs_mode, uint8_t *gs_out, float *coords3d);
// a13/a14 + a15 (offset 1) fused: the cracky GS image only ever exists tile by tile in shared memory
int backproject_fill_device(rsdsfm_ctx *, const uint8_t *image, const double *depth, int layout, int rows, int cols,
                            const double *K4, const double *R, const double *t, int gs_mode, uint8_t *rectified);
// a15: interpolateCrackyImage (camera.cc:753-774)
int fill_cracks_device(rsdsfm_ctx *, const uint8_t *in, int rows, int cols, unsigned offset, uint8_t *out);

// meanReprojectionError / createErrorImage (camera.cc:503-691); poses24 = per scanline
// original R[9], t[3] and relocated R[9], t[3]; sums5 (device): scale sum, entries, outliers, error sum, points
int reproj_device(rsdsfm_ctx *, const float *est, const double *ux, const double *uy, const double *uz, const double *poses24,
                  const double *depth_est, int layout, int rows, int cols, const double *K4, double max_norm, float *truep,
                  uint8_t *error_image, double *gt_depth, double *sums5);
// calculateTrueFlow (camera.cc:209-249); poses2 = frame 2's scanlines, rows x 12 (R[9], t[3])
int true_flow_device(rsdsfm_ctx *, const double *ux, const double *uy, const double *uz, const double *poses2, int layout,
                     int rows, int cols, const double *K4, double *flow);

}  // namespace rsdsfm
