/*
 * rsdsfm.h -- C ABI of the B200-native dense optimisation core of RS-aware differential SfM.
 *
 * The reference (ThomasZiegler/RS-aware-differential-SfM) has no plugin / FFI boundary: the path
 * is reached through plain C++ functions and classes.  Each entry point
 * below names the reference interface it replaces (file:line relative to the reference's src/).
 * The thin C++ classes in rs-aware-differential-sfm_b200/host/ carry the reference's names and
 * signatures and forward here; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - POD only, caller-owned buffers, `int` status return (RSDSFM_OK == 0), no exceptions.
 *   - One rsdsfm_ctx per GPU (and CUDA stream); contexts are independent and may be used from
 *     different threads; a single context is not re-entrant.
 *   - `mem` says where the ARRAY arguments of a call live: RSDSFM_HOST (the library stages them
 *     through device scratch with cudaMemcpyAsync on the context's stream) or RSDSFM_DEVICE
 *     (device pointers, nothing is copied; the call is asynchronous on the context's stream
 *     unless a scalar result has to be returned to the host).  Small fixed-size vectors
 *     (v[3], w[3], K4[4], k, summaries) are always host memory.
 *   - Layouts follow the reference's Eigen / OpenCV types: Array2Xd = interleaved pairs,
 *     Array3Xd = interleaved triples, MatrixXd(rows,cols) = column-major, cv::Mat 8UC3 =
 *     row-major interleaved BGR, cv::Mat_<Point2d> flow = row-major interleaved (dx,dy) doubles.
 *   - There is NO CPU fallback: every compute entry point fails with RSDSFM_ERR_CUDA when no
 *     sm_100 device / driver is usable.
 */
#ifndef RSDSFM_H
#define RSDSFM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSDSFM_VERSION 100

#if defined(__GNUC__)
#define RSDSFM_API __attribute__((visibility("default")))
#else
#define RSDSFM_API
#endif

enum {
    RSDSFM_OK = 0,
    RSDSFM_ERR_CUDA = 1,        /* CUDA runtime / driver error (see rsdsfm_last_error) */
    RSDSFM_ERR_ARG = 2,         /* invalid argument */
    RSDSFM_ERR_NOMEM = 3,       /* device allocation failed */
    RSDSFM_ERR_INTERNAL = 4     /* in-kernel watchdog or consistency check tripped */
};

enum { RSDSFM_HOST = 0, RSDSFM_DEVICE = 1 };
enum { RSDSFM_DEPTH_COLMAJOR = 0 /* Eigen MatrixXd, (y,x) at y + x*rows */, RSDSFM_DEPTH_ROWMAJOR = 1 };

typedef struct rsdsfm_ctx rsdsfm_ctx;

/* Solver options = the Ceres 1.14 Solver::Options fields the reference leaves at their defaults
 * (nonlinearRefinement.cc:159-163, :224-228 set only linear_solver_type = DENSE_SCHUR). */
typedef struct {
    int max_num_iterations;             /* 50    */
    double function_tolerance;          /* 1e-6  */
    double gradient_tolerance;          /* 1e-10 */
    double parameter_tolerance;         /* 1e-8  */
    double initial_trust_region_radius; /* 1e4   */
    double max_trust_region_radius;     /* 1e16  */
    double min_trust_region_radius;     /* 1e-32 */
    double min_relative_decrease;       /* 1e-3  */
    double min_lm_diagonal;             /* 1e-6  */
    double max_lm_diagonal;             /* 1e32  */
    int max_num_consecutive_invalid_steps; /* 5  */
} rsdsfm_lm_options;

enum { RSDSFM_CONVERGENCE = 0, RSDSFM_NO_CONVERGENCE = 1, RSDSFM_FAILURE = 2 };
enum {
    RSDSFM_REASON_NONE = 0, RSDSFM_REASON_PARAMETER_TOL = 1, RSDSFM_REASON_FUNCTION_TOL = 2,
    RSDSFM_REASON_GRADIENT_TOL = 3, RSDSFM_REASON_MAX_ITER = 4, RSDSFM_REASON_MIN_RADIUS = 5,
    RSDSFM_REASON_INVALID_STEPS = 6, RSDSFM_REASON_EVAL_FAILED = 7, RSDSFM_REASON_NONFINITE_INPUT = 8
};

/* What ceres::Solver::Summary::BriefReport() / total_time_in_seconds gave the reference
 * (nonlinearRefinement.cc:165-169, :230-234). */
typedef struct {
    int termination;          /* RSDSFM_CONVERGENCE / NO_CONVERGENCE / FAILURE */
    int reason;
    int iterations;           /* trust-region steps computed (incl. a final discarded one) */
    int num_successful;
    int num_unsuccessful;
    double initial_cost;
    double final_cost;
    double final_radius;
    double final_gradient_max_norm;
    double device_ms;         /* GPU time of the solve (CUDA events on the context's stream) */
} rsdsfm_lm_summary;

/* ---- context ------------------------------------------------------------------------- */
RSDSFM_API int rsdsfm_version(void);
RSDSFM_API void rsdsfm_lm_default_options(rsdsfm_lm_options *opts);
/* device: CUDA ordinal.  stream: a cudaStream_t to run on (e.g. the caller's framework stream),
 * or NULL to let the context create its own non-blocking stream. */
RSDSFM_API int rsdsfm_create(int device, void *cuda_stream, rsdsfm_ctx **out);
RSDSFM_API void rsdsfm_destroy(rsdsfm_ctx *ctx);
RSDSFM_API const char *rsdsfm_last_error(rsdsfm_ctx *ctx);   /* ctx may be NULL: error of a failed create */
RSDSFM_API int rsdsfm_synchronize(rsdsfm_ctx *ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
RSDSFM_API long long rsdsfm_launch_count(rsdsfm_ctx *ctx);
/* Per-kernel timing of the LM passes with CUDA events on the context's stream (what Ceres'
 * Summary::FullReport timing breakdown gave the reference user).  enable resets the counters.
 * out8: [0] pass-A ms, [1] pass-A launches, [2] pass-A residual blocks processed,
 *       [3] pass-B ms, [4] pass-B launches, [5] pass-B residual blocks processed, [6..7] reserved. */
RSDSFM_API int rsdsfm_profile_enable(rsdsfm_ctx *ctx, int on);
RSDSFM_API int rsdsfm_profile_read(rsdsfm_ctx *ctx, double *out8);
/* out8 (ms, accumulated like above): pass-A pixel loop, CTA reduction, grid reduction + controller;
 * the same three for pass B; then the controller logic alone (after the row reduction) for pass A
 * and pass B -- measured by CTA 0 / the controller CTA with %globaltimer. */
RSDSFM_API int rsdsfm_profile_detail(rsdsfm_ctx *ctx, double *out8);

/* Page-locked host buffers for the RSDSFM_HOST paths (asynchronous copies need them; the pipelined
 * sequence entry points overlap nothing with pageable memory).  write_combined: for buffers the CPU
 * only writes and the GPU only reads (inputs) -- uncached on the CPU side, no snoop on the DMA read. */
RSDSFM_API int rsdsfm_host_alloc(size_t bytes, int write_combined, void **out);
RSDSFM_API int rsdsfm_host_free(void *p);

/* ---- row split of ONE solve over several GPUs (BASELINE.json config 4; the reference is single-threaded, main.cc
 * has nothing to replace here) ---------------------------------------------------------------------------------
 * A group of contexts (one per GPU; one process per GPU, or several contexts of one process) solves one
 * nonLinearRefinement problem together: every member passes ITS share of the residual blocks (any contiguous split of
 * the consensus set) to rsdsfm_refine / rsdsfm_estimate_inverse_depths, which become collective calls -- every member
 * must make the same sequence of calls.  Each member receives the complete refined motion, the same termination
 * report, and the depths of its share.  Per LM iteration one row of 84 sums crosses the GPUs, through peer memory
 * (NVLink), from inside the persistent kernels; there is no NCCL call and no host round trip.
 *   1. every member: rsdsfm_peer_export -> a 64-byte CUDA IPC handle of its mailbox;
 *   2. the caller exchanges the handles (e.g. torch.distributed.all_gather) and every member calls
 *      rsdsfm_peer_connect with all of them, in group order;
 *   3. a barrier of the caller's choice (the mailboxes are cleared by connect), then the collective solves.
 * rsdsfm_peer_connect_local does 1-2 for contexts of ONE process.  rsdsfm_peer_disconnect returns a context to
 * single-GPU operation.  At most 8 members. */
#define RSDSFM_PEER_HANDLE_BYTES 64
RSDSFM_API int rsdsfm_peer_export(rsdsfm_ctx *ctx, void *handle_out);
RSDSFM_API int rsdsfm_peer_connect(rsdsfm_ctx *ctx, int n_members, int my_index, const void *handles);
RSDSFM_API int rsdsfm_peer_connect_local(rsdsfm_ctx **members, int n_members);
RSDSFM_API int rsdsfm_peer_disconnect(rsdsfm_ctx *ctx);

/* ---- a2: flatten + normalise glue (main.cc:398-432, errorMeasure.cpp:66-97) --------------- */
/* flow_img: rows*cols*2.  Outputs (each 2*rows*cols doubles) are pre-filled like the reference
 * (coord = 1, flow = 0) and the kept pixels (|flow|^2 > flow_threshold) are compacted in
 * COLUMN-MAJOR pixel order.  pixel_index (nullable, rows*cols int32): pixel_index[i] =
 * col*rows + row of kept point i.  *n_out = number of kept points (host). */
RSDSFM_API int rsdsfm_flatten(rsdsfm_ctx *ctx, int mem, const double *flow_img, int rows, int cols, const double *K4,
                   double gamma, double flow_threshold, double *coord, double *flow, double *coord_px,
                   double *flow_px, int32_t *pixel_index, int *n_out);

/* ---- a3/a4: minimal::getAlpha / getAlphaK (minimal.cc:179-197) ---------------------------- */
RSDSFM_API int rsdsfm_alpha(rsdsfm_ctx *ctx, int mem, const double *flow_px, const double *q_px, int n, double h,
                 double gamma, double *alpha, double *alpha_k);

/* ---- a5: minimal::calculateVelocities (minimal.cc:36-177) -------------------------------- */
/* Host computation (tiny dense LA, no context needed).  q9,u9: 2x9 interleaved; out7 = w,v,k. */
RSDSFM_API int rsdsfm_solve9(const double *q9, const double *u9, const double *alpha9, const double *alpha_k9,
                  int use_alpha_k, double *out7);

/* ---- a6 + a8: minimal::ransac scoring with an injected hypothesis list -------------------- */
/* (minimal.cc:209-306 with nonlinearRefinement.cc:109-180 inside.)  hyps: H x 7 (w,v,k), host
 * memory.  For every hypothesis the inverse depth of every point is estimated by the
 * Ceres-equivalent LM of estimateInverseDepths and inliers are counted (error < tolerance).
 * Outputs: counts[H], sumerr[H] (host); *best_idx by the reference's rule (more inliers, or as
 * many and a smaller error sum); mask_best[n] (uint8, 1 = inlier) and inv_depth_best[n] of the
 * winner live in `mem`. */
RSDSFM_API int rsdsfm_ransac_score(rsdsfm_ctx *ctx, int mem, const double *q, const double *u, const double *alpha,
                        const double *alpha_k, int n, const double *hyps, int H, double tolerance,
                        int *counts, double *sumerr, int *best_idx, uint8_t *mask_best,
                        double *inv_depth_best);

/* minimal::ransac with an injected SAMPLE list: samples = H x 9 point indices (host).  Runs the
 * 9-point solver on each sample, then rsdsfm_ransac_score.  hyps_out (nullable, host): H x 7. */
RSDSFM_API int rsdsfm_ransac(rsdsfm_ctx *ctx, int mem, const double *q, const double *u, const double *alpha,
                  const double *alpha_k, int n, int use_alpha_k, const int32_t *samples, int H,
                  double tolerance, int *counts, double *sumerr, int *best_idx, double *best7,
                  uint8_t *mask_best, double *inv_depth_best, double *hyps_out);

/* tail of minimal::ransac (minimal.cc:291-305): consensus set in ascending index order.
 * inliers3: 3 x m (x, y, z = 1/inv_depth); *m_out on the host. */
RSDSFM_API int rsdsfm_gather_inliers(rsdsfm_ctx *ctx, int mem, const double *q, const double *alpha,
                          const double *alpha_k, int n, const uint8_t *mask, const double *inv_depth,
                          double *inliers3, double *alpha_in, double *alpha_k_in, int32_t *index_in,
                          int *m_out);

/* ---- a8: nonlinear_refinement::estimateInverseDepths (nonlinearRefinement.cc:109-180) ----- */
RSDSFM_API int rsdsfm_estimate_inverse_depths(rsdsfm_ctx *ctx, int mem, const double *coord, const double *flow, int n,
                                   const double *v, const double *w, double k, const double *alpha,
                                   const double *alpha_k, double *inv_depth, rsdsfm_lm_summary *summary);

/* ---- a9: nonlinear_refinement::nonLinearRefinement (nonlinearRefinement.cc:183-252) ------- */
/* flow: the array the caller passes (2 x >= m; residual i reads flow(:,i) of it exactly like the
 * reference, Q1) unless flow_index (nullable, m int32) gives the repaired pairing.
 * inliers3: 3 x m (x, y, z).  v, w, k: host, in/out.  z_out[m]: refined depths (1/d).
 * opts may be NULL (Ceres defaults). */
RSDSFM_API int rsdsfm_refine(rsdsfm_ctx *ctx, int mem, const double *flow, const double *inliers3, const double *alpha,
                  const double *alpha_k, int m, double *v, double *w, double *k, int const_acceleration,
                  const int32_t *flow_index, const rsdsfm_lm_options *opts, double *z_out,
                  rsdsfm_lm_summary *summary);

/* ---- a10: sign fix + depth raster glue (main.cc:466-509, errorMeasure.cpp:162-210) -------- */
/* inliers3 z row and v are sign-fixed in place; depth_map (rows*cols doubles, `layout`) is zero
 * filled and rasterised; depth_img (nullable, rows*cols uint8 row-major) is the 8-bit image. */
RSDSFM_API int rsdsfm_depth_glue(rsdsfm_ctx *ctx, int mem, double *inliers3, int m, double *v, const double *K4,
                      int rows, int cols, double z_min_init, int layout, double *depth_map,
                      uint8_t *depth_img);

/* ---- a12: Camera::setPose -> RsFrame::setRelativePose (camera.cc:340, rsframe.cc:771-800) - */
/* Host computation.  R: rows x 9 (row-major 3x3), t: rows x 3. */
RSDSFM_API int rsdsfm_set_relative_pose(const double *v, const double *w, double k, double gamma, int rows, double *R,
                             double *t);

/* ---- a13 + a14: Camera::backProject / backProjectGs (rsframe.cc:629-736, 803-878) -------- */
/* image: rows*cols*3 BGR; depth: rows*cols doubles in `layout`; R,t: per-scanline relative poses
 * (host, rows x 9 / rows x 3).  gs_out: rows*cols*3; coords3d (nullable): rows*cols*3 float. */
RSDSFM_API int rsdsfm_backproject(rsdsfm_ctx *ctx, int mem, const uint8_t *image, const double *depth, int layout,
                       int rows, int cols, const double *K4, const double *R, const double *t, int gs_mode,
                       uint8_t *gs_out, float *coords3d);

/* ---- a15: Camera::interpolateCrackyImage (camera.cc:753-774) ------------------------------ */
RSDSFM_API int rsdsfm_fill_cracks(rsdsfm_ctx *ctx, int mem, const uint8_t *in, int rows, int cols, unsigned offset,
                       uint8_t *out);

/* ---- accuracy metric of the rectified geometry ------------------------------ */
/* RsFrame::relocatePose (rsframe.cc:953-967), host computation: scanline 0 keeps its pose, for
 * i >= 1  t_i -= t_0,  R_i = R_0^-1 R_i.  R: rows x 9 (row-major 3x3), t: rows x 3. */
RSDSFM_API int rsdsfm_relocate_pose(const double *R, const double *t, int rows, double *R_out, double *t_out);

/* Camera::meanReprojectionError (camera.cc:593-691) and, when error_image != NULL,
 * Camera::createErrorImage (camera.cc:503-590), including RsFrame::getGroundtruthDepthMap
 * (rsframe.cc:416-436) and relocatePose.
 *   coords3d   rows*cols*3 float: RsFrame::get3dCoordinates() as left by rsdsfm_backproject
 *   unproj_*   rows*cols doubles each: the frame's unprojection maps (world point per RS pixel)
 *   R_gt,t_gt  host, rows x 9 / rows x 3: ground-truth camera-from-world pose of every scanline
 *   depth_est  rows*cols: the frame's depth map (used where the ground-truth depth is exactly 0,
 *              like planeToSpace's default argument, rsframe.cc:657-659)
 *   layout     RSDSFM_DEPTH_* of unproj_*, depth_est and gt_depth_map
 * Outputs (host scalars): *mean_error; nullable *mean_scale, *num_outliers, *points_used;
 * nullable arrays in `mem`: error_image (rows*cols, row-major 8-bit, max_norm = pixel value 255)
 * and gt_depth_map (rows*cols). */
RSDSFM_API int rsdsfm_reprojection_error(rsdsfm_ctx *ctx, int mem, const float *coords3d, const double *unproj_x,
                          const double *unproj_y, const double *unproj_z, const double *R_gt, const double *t_gt,
                          const double *depth_est, int layout, int rows, int cols, const double *K4,
                          double max_norm, double *mean_error, double *mean_scale, int *num_outliers,
                          int *points_used, uint8_t *error_image, double *gt_depth_map);

/* ---- ground-truth flow between two synthetic RS frames ---------------------- */
/* Camera::calculateTrueFlow (camera.cc:209-249) with RsFrame::calculateImageCoordinatesRsFrame
 * (rsframe.cc:740-768): the world point of every pixel of frame 1 (unproj_*, rows*cols doubles in
 * `layout`) is projected with every scanline pose of frame 2 (R2: rows x 9, t2: rows x 3, host) and
 * the pose whose row is closest to the projected y is used.  flow: rows*cols*2 row-major (dx,dy),
 * i.e. the cv::Mat_<Point2d> that rsdsfm_flatten consumes. */
RSDSFM_API int rsdsfm_true_flow(rsdsfm_ctx *ctx, int mem, const double *unproj_x, const double *unproj_y,
                          const double *unproj_z, const double *R2, const double *t2, int layout, int rows,
                          int cols, const double *K4, double *flow);

/* ---- fused driver of the timed region "refine + rectify" (main.cc:457-523) ---------------- */
/* nonLinearRefinement -> sign fix -> depth raster -> setPose -> backProject(Gs) ->
 * interpolateCrackyImage(.,1) for one frame pair, without leaving the device.
 * Array arguments live in `mem`; v,w,k (in/out), K4, summary on the host.
 * Outputs: z_out[m] (sign-fixed refined depths), depth_map (rows*cols, `layout`),
 * rectified (rows*cols*3). */
RSDSFM_API int rsdsfm_refine_rectify(rsdsfm_ctx *ctx, int mem, const double *flow, const double *inliers3,
                          const double *alpha, const double *alpha_k, int m, double *v, double *w, double *k,
                          int const_acceleration, int gs_mode, const uint8_t *image, int rows, int cols,
                          const double *K4, double gamma, int layout, double *z_out, double *depth_map,
                          uint8_t *rectified, rsdsfm_lm_summary *summary);

/* ---- the same step over a sequence of frame pairs, software-pipelined ---------------------- */
/* One entry per frame pair: the arguments of rsdsfm_refine_rectify that change from pair to
 * pair.  v,w,k are in/out (start motion = the RANSAC winner; refined and sign-fixed on return). */
typedef struct rsdsfm_pair_io {
    const double *flow, *inliers3, *alpha, *alpha_k;   /* as in rsdsfm_refine_rectify */
    const uint8_t *image;                              /* rows*cols*3 */
    int m;
    int status;                                        /* out: RSDSFM_OK or the pair's error code */
    double v[3], w[3], k;
    double *z_out, *depth_map;                         /* m, rows*cols */
    uint8_t *rectified;                                /* rows*cols*3 */
    rsdsfm_lm_summary summary;                         /* out */
} rsdsfm_pair_io;

/* The reference handles one frame pair per run (evaluateSingleRun, main.cc:302-560; the sweep in
 * main.cc:148-300 and errorMeasure.cpp:99-226 repeat it); pairs are independent, so this entry point keeps
 * several of them in flight on one GPU.  Every pair runs on a compute lane -- an internal context with its own
 * stream and buffers whose LM solve takes a quarter of the SMs (half of them for sequences of fewer than 12
 * pairs) -- and goes to whichever lane is free first: one solve's grid exchanges and serial controller steps
 * are covered by the other solves' pixel sweeps, and with mem = RSDSFM_HOST the uploads (one copy stream, pair
 * order) and downloads (a copy stream per lane) of some pairs overlap the compute of others (use pinned host
 * memory for the copies to overlap).  Up to 16 lanes (10 with device buffers); their buffers are sized for
 * rows * cols on first use and kept.  Environment: RSDSFM_ACTIVE_LANES (solves sharing the SMs; 1 = full-GPU solves), RSDSFM_LANES,
 * RSDSFM_TRACE=1 (per-solve device time stamps and per-call host times on stderr).
 * Results do not depend on the lanes: the LM kernel sums over fixed strips of residual blocks whatever its
 * grid (csrc/lm_kernel.cuh, kStrips), so every pair comes out bit-identical to a single rsdsfm_refine_rectify
 * call.  Returns the first failing pair's code (all pairs are attempted; see rsdsfm_pair_io.status). */
RSDSFM_API int rsdsfm_refine_rectify_sequence(rsdsfm_ctx *ctx, int mem, int n_pairs, rsdsfm_pair_io *pairs,
                          int const_acceleration, int gs_mode, int rows, int cols, const double *K4,
                          double gamma, int layout);

/* ---- the same step fed with what minimal::ransac returned: the compact host interface --------- */
/* The inputs of the step are the cached flow field, the RS frame and the RANSAC result.  The expanded arrays
 * the reference passes around between main.cc:398 and :457 (normalised coordinates, gamma-scaled flow,
 * alpha / alpha_k, the consensus set as x, y, z triples) are functions of the flow field and of the winner's
 * consensus mask and inverse depths, so this entry point takes those -- exactly the outputs of rsdsfm_ransac --
 * and rebuilds the rest on the device, bit-identical to the stage-wise calls (same residual pairing as
 * nonlinearRefinement.cc:209-216).  Per 1080p pair 58 MB cross the host link upwards instead of 122 MB
 * (42 MB with float32 flow), and 23 MB come back instead of 39 MB when the depth map is not requested. */
typedef struct rsdsfm_compact_pair_io {
    const void *flow_img;          /* rows*cols*2 (dx, dy) row-major: doubles, or floats with flow_f32 (DeepFlow's
                                    * output is float32 widened to double, camera.cc:262-274: lossless) */
    const uint8_t *image;          /* rows*cols*3 */
    const uint8_t *mask;           /* n: consensus mask over the flattened points (rsdsfm_ransac mask_best) */
    const double *inv_depth;       /* n: the winner's inverse depths (rsdsfm_ransac inv_depth_best) */
    int n, m;                      /* flattened points (rsdsfm_flatten *n_out), consensus-set size (number of mask
                                    * entries set); checked on the device: a mismatch fails the pair with RSDSFM_ERR_ARG */
    int status;                    /* out */
    double v[3], w[3], k;          /* in: the RANSAC winner; out: refined, sign-fixed */
    double *z_out;                 /* m: refined, sign-fixed depths in consensus-set order */
    double *depth_map;             /* rows*cols, or NULL: not wanted */
    uint8_t *rectified;            /* rows*cols*3 */
    rsdsfm_lm_summary summary;     /* out */
} rsdsfm_compact_pair_io;

/* Pipelined over compute lanes like rsdsfm_refine_rectify_sequence (bit-identical to single calls).
 * flow_threshold: the |flow|^2 cut of the flattening (1e-10 in the reference). */
RSDSFM_API int rsdsfm_refine_rectify_compact_sequence(rsdsfm_ctx *ctx, int mem, int n_pairs, rsdsfm_compact_pair_io *pairs,
                          int flow_f32, double flow_threshold, int const_acceleration, int gs_mode, int rows, int cols,
                          const double *K4, double gamma, int layout);

/* ---- a2..a15 in one call: evaluateSingleRun's compute (main.cc:398-523) --------------------- */
typedef struct rsdsfm_pipeline_params {
    int rows, cols;
    double K4[4];               /* fx, fy, cx, cy */
    double gamma;               /* readout ratio (main.cc:321) */
    double flow_threshold;      /* 1e-10 in the reference (main.cc:311) */
    double ransac_tolerance;    /* main.cc:310 */
    int num_hypotheses;         /* RANSAC trials H (main.cc:304) */
    int const_acceleration;     /* use_acceleration_mode */
    int gs_mode;                /* use_global_shutter_mode: alpha := 1 (main.cc:441-444), backProjectGs */
    int use_refinement;         /* main.cc:457 */
    int repair_pairing;         /* 0 = reference behaviour (residual i reads flow(:, i));
                                 * 1 = residual i reads the flow of inlier i */
    int layout;                 /* depth_map layout, RSDSFM_DEPTH_* */
    int emulate_padding;        /* 0 = the point set is the n kept flow vectors (errorMeasure.cpp:96-97 truncates);
                                 * 1 = main.cc:398-447 as written: the arrays stay rows*cols long and the tail
                                 *     (coord = (1,1), flow = 0, alpha = 1) is sampled, scored and can join the
                                 *     consensus set.  `samples` / `draws` then index / are reduced
                                 *     modulo rows*cols points; the out-of-image raster writes of the reference
                                 *     (main.cc:501-508) are dropped. */
} rsdsfm_pipeline_params;

typedef struct rsdsfm_pipeline_io {
    const double *flow_img;     /* rows*cols*2 (dx, dy) row-major, in `mem` */
    const uint8_t *image;       /* rows*cols*3 BGR, in `mem` */
    const int32_t *samples;     /* host, H*9 indices into the flattened order, or NULL */
    const uint32_t *draws;      /* host, H*9 raw rand() values mapped to indices exactly like
                                 * minimal.cc:226-244 (rand() % n_temp on a persistent index vector of the
                                 * point set: n kept vectors, or rows*cols with emulate_padding);
                                 * used when samples == NULL */
    double *depth_map;          /* out, rows*cols, in `mem` */
    uint8_t *rectified;         /* out, rows*cols*3, in `mem` */
    /* results (host) */
    int status;
    int n, m, best_idx;         /* valid flow vectors, consensus-set size, winning trial */
    double ransac_motion[7];    /* v, w, k of the winning hypothesis */
    double v[3], w[3], k;       /* final motion (refined when use_refinement), sign-fixed */
    rsdsfm_lm_summary summary;
} rsdsfm_pipeline_io;

/* flatten -> alpha -> RANSAC (fit + score + gather) -> [refine] -> sign fix -> depth raster ->
 * setPose -> backProject(Gs) -> interpolateCrackyImage, intermediates never leave the device. */
RSDSFM_API int rsdsfm_pipeline_pair(rsdsfm_ctx *ctx, int mem, const rsdsfm_pipeline_params *params,
                          rsdsfm_pipeline_io *io);
/* The same over a sequence; with RSDSFM_HOST buffers the next pair's flow/image upload and the
 * previous pair's download overlap the current pair's compute. */
RSDSFM_API int rsdsfm_pipeline_sequence(rsdsfm_ctx *ctx, int mem, const rsdsfm_pipeline_params *params, int n_pairs,
                          rsdsfm_pipeline_io *pairs);

#ifdef __cplusplus
}
#endif
#endif /* RSDSFM_H */
