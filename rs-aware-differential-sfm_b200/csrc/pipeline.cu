// pipeline.cu -- the fused drivers behind include/rsdsfm.h:
//   rsdsfm_refine_rectify            main.cc:457-523 for one frame pair, one host synchronisation
//   rsdsfm_refine_rectify_sequence   the same over a sequence: several pairs in flight on compute lanes
//                                    (uploads | solves on a quarter of the SMs each | downloads overlap)
//   rsdsfm_pipeline_pair / _sequence main.cc:398-523 (flatten .. crack fill) without leaving the device
// Everything here composes the stage functions of stages.h; there is no arithmetic in this file
// apart from the reference's sample draw (minimal.cc:226-244) on the host.
#include <chrono>
#include <thread>
#include <unordered_map>
#include <vector>

#include "stages.h"

namespace rsdsfm {

// ---- one "refine + rectify" step on device pointers --------------------------------------------
struct StepArgs {
    const double *flow, *inliers3, *alpha, *alpha_k;
    const int32_t *flow_index;
    const uint8_t *image;
    int m, const_acc, gs_mode, rows, cols, layout;
    const double *K4;
    double gamma;
    double *z, *depth_map;
    uint8_t *rectified;
    // compact inputs (rsdsfm_refine_rectify_compact*): set flow_img and leave flow .. alpha_k NULL
    const void *flow_img = nullptr;
    int flow_f32 = 0;
    const uint8_t *mask = nullptr;
    const double *inv_depth = nullptr;
    int n = 0;
    double flow_threshold = 0.0;
};

// Queues the whole step on ctx->stream, including the small read-backs (LM control block, depth
// statistics) into the current I/O slot's pinned area.  No synchronisation.
static int queue_step(rsdsfm_ctx *ctx, const StepArgs &a, const double *v, const double *w, double k)
{
    const size_t tot = (size_t)a.rows * a.cols;
    RS_TRY(ensure(ctx, ctx->misc, 256));
    RS_TRY(ensure(ctx, ctx->poses, sizeof(double) * 12 * (size_t)a.rows));
    double *stats = (double *)ctx->misc.p;
    double *dR = (double *)ctx->poses.p, *dt = dR + 9 * (size_t)a.rows;
    // nonLinearRefinement (main.cc:457): the refined motion stays in the solver's device control
    // block and feeds the pose kernel directly
    // (the solve's epilogue also leaves the per-CTA sums of z the sign fix needs)
    RS_TRY(ensure(ctx, ctx->sums, sizeof(double) * 3 * (size_t)ctx->num_sms));
    double *zrows = (double *)ctx->sums.p;
    const double *xyz = a.inliers3;
    int xs = 3;
    if (a.flow_img) {
        // compact inputs: coordinates, alpha factors, pairing and start depths are rebuilt on the device, straight
        // into the solver's layout (no expanded arrays ever exist)
        const size_t mm = (size_t)(a.m > 0 ? a.m : 1);
        RS_TRY(ensure(ctx, ctx->pipe[14], sizeof(double) * mm));          // start depths z (restored on FAILURE)
        RS_TRY(ensure(ctx, ctx->pipe[15], sizeof(double) * 2 * mm));      // normalised coordinates of the inliers
        void *blk = nullptr; double *d0 = nullptr; int *flag = nullptr;
        RS_TRY(lm_input_buffers(ctx, a.m, &blk, &d0, &flag));
        RS_TRY(compact_build_device(ctx, a.flow_img, a.flow_f32, a.rows, a.cols, a.K4, a.gamma, a.flow_threshold, a.mask, a.inv_depth,
                                    a.n, a.m, blk, d0, (double *)ctx->pipe[14].p, (double *)ctx->pipe[15].p, flag));
        RS_TRY(refine_prepared_async(ctx, a.m, v, w, k, a.const_acc, nullptr, (const double *)ctx->pipe[14].p, a.z, zrows));
        xyz = (const double *)ctx->pipe[15].p;
        xs = 2;
    } else {
        RS_TRY(refine_async(ctx, a.flow, a.inliers3, a.alpha, a.alpha_k, a.m, v, w, k, a.const_acc, a.flow_index, nullptr, a.z, zrows));
    }
    // sign fix + depth raster (main.cc:466-509)
    RS_TRY(glue_device(ctx, a.z, 1, xyz, xs, a.m, a.K4, a.rows, a.cols, INFINITY, a.layout, a.depth_map, nullptr, stats,
                       zrows, lm_grid_size(ctx)));
    // setPose (main.cc:516) -> per-scanline poses, with the sign-fixed v
    RS_TRY(poses_device(ctx, lm_motion_device(ctx), stats, a.gamma, a.rows, dR, dt));
    // backProject / backProjectGs (main.cc:518-522) + interpolateCrackyImage (main.cc:523)
    RS_TRY(backproject_fill_device(ctx, a.image, a.depth_map, a.layout, a.rows, a.cols, a.K4, dR, dt, a.gs_mode, a.rectified));
    return lm_collect_enqueue(ctx, stats);
}

// Parses the current I/O slot's read-backs (which must have completed).
static int finish_step(rsdsfm_ctx *ctx, int nf, int m, double *v, double *w, double *k, rsdsfm_lm_summary *summary,
                       bool *overflow)
{
    Motion mot;
    for (int j = 0; j < 3; ++j) { mot.v[j] = v[j]; mot.w[j] = w[j]; }
    mot.k = *k;
    RS_TRY(lm_collect_finish(ctx, nf, m, &mot, summary, overflow));
    if (*overflow) return RSDSFM_OK;
    const double sign = (pinned_stats(ctx)[3] < 0) ? -1.0 : 1.0;      // z_mean < 0: v *= -1 (main.cc:475-478)
    for (int j = 0; j < 3; ++j) { v[j] = mot.v[j] * sign; w[j] = mot.w[j]; }
    *k = mot.k;
    return RSDSFM_OK;
}

static int ensure_io(rsdsfm_ctx *ctx)
{
    if (ctx->s_in) return RSDSFM_OK;
    RS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    RS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    for (int j = 0; j < 2; ++j) {
        RS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_in[j], cudaEventDisableTiming));
        RS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_cdone[j], cudaEventDisableTiming));
        RS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_out[j], cudaEventDisableTiming));
    }
    return RSDSFM_OK;
}

static void drain(rsdsfm_ctx *ctx)
{
    if (ctx->s_in) cudaStreamSynchronize(ctx->s_in);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->s_out) cudaStreamSynchronize(ctx->s_out);
}

// ---- one frame pair of either host interface, as the sequence driver sees it -----------------------
// in[0..4]: expanded inputs flow, inliers3, alpha, alpha_k, image  |  compact inputs flow_img, mask, inv_depth, -, image
struct SeqPair {
    const void *in[5];
    size_t in_bytes[5];
    int m, n;
    double *v, *w, *k;
    double *z_out, *depth_map;           // depth_map may be NULL with the compact interface (not downloaded / not kept)
    uint8_t *rectified;
    rsdsfm_lm_summary *summary;
    int *status;
};
struct SeqCommon {
    int compact, flow_f32;
    double flow_threshold;
    int const_acc, gs_mode, rows, cols, layout;
    const double *K4;
    double gamma;
};

static StepArgs step_args(const SeqCommon &C, const SeqPair &p, const void *const in[5], double *z, double *depth_map, uint8_t *rectified)
{
    StepArgs a{};
    a.image = (const uint8_t *)in[4];
    a.m = p.m; a.const_acc = C.const_acc; a.gs_mode = C.gs_mode; a.rows = C.rows; a.cols = C.cols; a.layout = C.layout;
    a.K4 = C.K4; a.gamma = C.gamma;
    a.z = z; a.depth_map = depth_map; a.rectified = rectified;
    if (C.compact) {
        a.flow_img = in[0]; a.flow_f32 = C.flow_f32; a.mask = (const uint8_t *)in[1]; a.inv_depth = (const double *)in[2];
        a.n = p.n; a.flow_threshold = C.flow_threshold;
    } else {
        a.flow = (const double *)in[0]; a.inliers3 = (const double *)in[1]; a.alpha = (const double *)in[2]; a.alpha_k = (const double *)in[3];
    }
    return a;
}

// Synchronous step with the caller's buffers in `mem` (I/O slot `slot`).
static int step_sync(rsdsfm_ctx *ctx, int mem, int slot, const SeqCommon &C, const SeqPair &p)
{
    const size_t mm = (size_t)p.m, tot = (size_t)C.rows * C.cols;
    const int b = 8 * slot;
    ctx->io_slot = slot;
    const void *d_in[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    void *d_z = nullptr, *d_dm = nullptr, *d_out = nullptr;
    for (int j = 0; j < 5; ++j) RS_TRY(stage_in(ctx, mem, b + j, p.in[j], p.in_bytes[j], &d_in[j]));
    RS_TRY(stage_out_reserve(ctx, mem, b + 5, p.z_out, sizeof(double) * mm, &d_z));
    if (p.depth_map) RS_TRY(stage_out_reserve(ctx, mem, b + 6, p.depth_map, sizeof(double) * tot, &d_dm));
    else { RS_TRY(ensure(ctx, ctx->stage[b + 6], sizeof(double) * tot)); d_dm = ctx->stage[b + 6].p; }   // the splat still reads it
    RS_TRY(stage_out_reserve(ctx, mem, b + 7, p.rectified, tot * 3, &d_out));
    rsdsfm_lm_summary local;
    rsdsfm_lm_summary *summary = p.summary ? p.summary : &local;
    memset(summary, 0, sizeof *summary);
    const StepArgs a = step_args(C, p, d_in, (double *)d_z, (double *)d_dm, (uint8_t *)d_out);
    for (int attempt = 0; attempt < 2; ++attempt) {
        RS_TRY(queue_step(ctx, a, p.v, p.w, *p.k));
        RS_TRY(stage_out(ctx, mem, p.z_out, d_z, sizeof(double) * mm));
        RS_TRY(stage_out(ctx, mem, p.depth_map, d_dm, sizeof(double) * tot));
        RS_TRY(stage_out(ctx, mem, p.rectified, d_out, tot * 3));
        RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));               // the one synchronisation of the step
        bool overflow = false;
        RS_TRY(finish_step(ctx, C.const_acc ? 7 : 6, p.m, p.v, p.w, p.k, summary, &overflow));
        if (!overflow) return RSDSFM_OK;                                 // else: exception list enlarged, run again
    }
    return fail(ctx, RSDSFM_ERR_INTERNAL, "refine_rectify: exception list overflow persisted");
}

static bool seq_pair_ok(const SeqCommon &C, const SeqPair &p)
{
    if (!(p.m > 0 && p.in[0] && p.in[1] && p.in[2] && p.in[4] && p.z_out && p.rectified)) return false;
    if (C.compact) return p.n >= p.m;
    return p.in[3] && p.depth_map;
}

// The sequence driver behind rsdsfm_refine_rectify_sequence and rsdsfm_refine_rectify_compact_sequence
// (see include/rsdsfm.h for the pipeline it implements).
static int sequence_core(rsdsfm_ctx *ctx, int mem, std::vector<SeqPair> &pairs, const SeqCommon &C)
{
    const int n_pairs = (int)pairs.size();
    if (n_pairs == 0) return RSDSFM_OK;
    RS_TRY(ensure_io(ctx));
    const size_t tot = (size_t)C.rows * C.cols;
    const int nf = C.const_acc ? 7 : 6;
    const bool host = (mem == RSDSFM_HOST);
    int first_err = RSDSFM_OK;
    int max_m = 0, n_ok = 0;
    size_t max_in[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < n_pairs; ++i) {
        SeqPair &p = pairs[i];
        *p.status = seq_pair_ok(C, p) ? RSDSFM_OK : RSDSFM_ERR_ARG;
        memset(p.summary, 0, sizeof *p.summary);
        if (*p.status == RSDSFM_OK) {
            ++n_ok;
            if (p.m > max_m) max_m = p.m;
            for (int j = 0; j < 5; ++j) if (p.in_bytes[j] > max_in[j]) max_in[j] = p.in_bytes[j];
        } else if (first_err == RSDSFM_OK)
            first_err = fail(ctx, RSDSFM_ERR_ARG, "refine_rectify sequence: bad pair argument");
    }
    if (max_m == 0) return first_err;
    // Buffers are sized for the frame, not for this call's pairs (m <= n <= rows * cols): a later sequence with a few
    // more inliers must not reallocate -- cudaFree / cudaMalloc of the lanes' buffers costs tens of milliseconds.
    if ((size_t)max_m < tot) max_m = (int)tot;
    {
        const size_t bound_c[5] = {0, tot, sizeof(double) * tot, 0, tot * 3};
        const size_t bound_x[5] = {sizeof(double) * 2 * tot, sizeof(double) * 3 * tot, sizeof(double) * tot, sizeof(double) * tot, tot * 3};
        for (int j = 0; j < 5; ++j) {
            const size_t bnd = C.compact ? bound_c[j] : bound_x[j];
            if (max_in[j] && max_in[j] < bnd) max_in[j] = bnd;
        }
    }

    // Compute lanes (see common.cuh): every pair runs on one lane, a context of its own whose LM solve takes
    // 1/active of the SMs.  A pair goes to whichever lane is free first (the iteration counts of the pairs of a
    // sequence differ by several times: a fixed rotation would leave lanes waiting behind the slowest pair).
    //   RSDSFM_ACTIVE_LANES  solves that share the SMs   (default 4, 2 for sequences of fewer than 12 pairs -- the last
    //                        solves of a sequence run with idle SMs beside them, the longer the more lanes there are;
    //                        1 = full-GPU solves, I/O still overlapped)
    //   RSDSFM_LANES         lanes                       (default 10 with device buffers, 16 with host buffers, where a
    //                        lane spends a third of its time in its copies -- measured 893 / 1000 / 1011 pairs/s end
    //                        to end with 8 / 12 / 16 lanes; fewer when 16 lanes' buffers would exceed ~8 GB)
    auto env_int = [](const char *name, int dflt) { const char *e = getenv(name); return (e && atoi(e) > 0) ? atoi(e) : dflt; };
    int active = env_int("RSDSFM_ACTIVE_LANES", n_ok >= 12 ? 4 : 2);
    if (getenv("RSDSFM_SINGLE_LANE")) active = 1;
    if (active > ctx->num_sms) active = ctx->num_sms;
    int dflt_lanes = host ? 16 : 10;
    while (dflt_lanes > 4 && (double)dflt_lanes * (double)tot * (host ? 140.0 : 100.0) > 8e9) dflt_lanes -= 2;   // bytes per pixel a lane holds
    int n_lanes = env_int("RSDSFM_LANES", dflt_lanes);
    if (n_lanes > kMaxLanes) n_lanes = kMaxLanes;
    if (n_lanes > n_ok) n_lanes = n_ok;
    if (ctx->n_peers > 1) n_lanes = 1;                      // row split: the peers step through the same solves in lockstep
    if (active > n_lanes) active = n_lanes;
    rsdsfm_ctx *lane[kMaxLanes] = {ctx};
    for (int l = 1; l < n_lanes; ++l) {
        if (!ctx->lanes[l - 1] && rsdsfm_create(ctx->device, nullptr, &ctx->lanes[l - 1]) != RSDSFM_OK)
            return fail(ctx, RSDSFM_ERR_CUDA, rsdsfm_last_error(nullptr));
        lane[l] = ctx->lanes[l - 1];
    }
    auto drain_all = [&]() { drain(ctx); for (auto *L : ctx->lanes) if (L) drain(L); };
    long long launches_before[kMaxLanes] = {0};
    auto restore = [&]() {                                   // lanes back to what single calls expect
        for (int l = 0; l < n_lanes; ++l) { lane[l]->lm_grid = 0; lane[l]->io_slot = 0; }
    };
    // size every buffer once, before anything is in flight
    const bool trace = getenv("RSDSFM_TRACE") != nullptr;
    const auto tc0 = std::chrono::steady_clock::now();
    drain_all();
    int rc0 = RSDSFM_OK;
    for (int l = 0; l < n_lanes && rc0 == RSDSFM_OK; ++l) {
        rsdsfm_ctx *L = lane[l];
        launches_before[l] = L->launches;
        L->lm_grid = active > 1 ? ctx->num_sms / active : 0;
        L->io_slot = 0;
        if (l > 0) { L->profile = ctx->profile; if (ctx->exc_cap > L->exc_cap) L->exc_cap = ctx->exc_cap; }
        rc0 = ensure_io(L);
        if (rc0 == RSDSFM_OK) rc0 = lm_reserve(L, max_m);
        const size_t sz[8] = {max_in[0], max_in[1], max_in[2], max_in[3], max_in[4], sizeof(double) * (size_t)max_m, sizeof(double) * tot, tot * 3};
        for (int j = (host ? 0 : 6); j < (host ? 8 : 7) && rc0 == RSDSFM_OK; ++j)       // device buffers: only the optional depth scratch
            if (sz[j]) rc0 = ensure(L, L->stage[j], sz[j]);
        if (rc0 == RSDSFM_OK && C.compact) {
            rc0 = ensure(L, L->pipe[14], sizeof(double) * (size_t)max_m);
            if (rc0 == RSDSFM_OK) rc0 = ensure(L, L->pipe[15], sizeof(double) * 2 * (size_t)max_m);
            if (rc0 == RSDSFM_OK) rc0 = ensure(L, L->flow_t, (C.flow_f32 ? sizeof(float) : sizeof(double)) * 2 * tot);
        }
        if (rc0 != RSDSFM_OK && L != ctx) ctx->err = L->err;
    }
    if (rc0 != RSDSFM_OK) { restore(); return rc0; }
    // the lanes have streams of their own: whatever the caller queued on the context's stream before this call
    // (e.g. the kernels that produced the device inputs) must be ordered before their work too
    RS_CUDA(ctx, cudaEventRecord(ctx->ev_in[1], ctx->stream));
    for (int l = 1; l < n_lanes; ++l) RS_CUDA(ctx, cudaStreamWaitEvent(lane[l]->stream, ctx->ev_in[1], 0));
    if (host) RS_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_in[1], 0));

    // A lane's life: upload (ctx->s_in, all uploads in pair order) -> compute (the lane's stream) -> download (the
    // lane's own s_out, so that a short pair's results do not queue behind a long pair's) -> parsed by the host.
    // The host only ever waits for "some lane has finished".
    int occupant[kMaxLanes];                     // pair on each lane, not yet finished
    for (int l = 0; l < kMaxLanes; ++l) occupant[l] = -1;
    auto finish = [&](int l) -> int {
        const int j = occupant[l];
        if (j < 0) return RSDSFM_OK;
        occupant[l] = -1;
        SeqPair &p = pairs[j];
        rsdsfm_ctx *L = lane[l];
        RS_CUDA(ctx, cudaEventSynchronize(L->ev_out[0]));
        bool overflow = false;
        int rc = finish_step(L, nf, p.m, p.v, p.w, p.k, p.summary, &overflow);
        if (rc == RSDSFM_OK && overflow) {
            // Exception list too small for this pair (it has been enlarged): let everything in flight complete
            // (the other lanes' read-backs stay where they are) and redo this pair synchronously on its lane.
            drain_all();
            rc = step_sync(L, mem, 0, C, p);
        }
        if (rc != RSDSFM_OK && L != ctx) ctx->err = L->err;
        *p.status = rc;
        return rc;
    };
    auto free_lane = [&]() -> int {              // a lane without occupant; else the first lane whose pair is complete
        for (int l = 0; l < n_lanes; ++l) if (occupant[l] < 0) return l;
        for (unsigned spin = 0;; ++spin) {
            for (int l = 0; l < n_lanes; ++l) {
                const cudaError_t q = cudaEventQuery(lane[l]->ev_out[0]);
                if (q != cudaErrorNotReady) return l;          // complete (or failed: finish() reports it)
            }
            if ((spin & 63u) == 63u) std::this_thread::yield();
        }
    };

    const auto tc1 = std::chrono::steady_clock::now();
    for (int i = 0; i < n_pairs; ++i) {
        SeqPair &p = pairs[i];
        if (*p.status != RSDSFM_OK) continue;
        const int l = free_lane();
        rsdsfm_ctx *L = lane[l];
        const size_t mm = (size_t)p.m;
        const int frc = finish(l);                // the lane's previous pair (complete, or about to be)
        if (frc != RSDSFM_OK && first_err == RSDSFM_OK) first_err = frc;
        int rc = [&]() -> int {
            const void *in[5] = {p.in[0], p.in[1], p.in[2], p.in[3], p.in[4]};
            double *z = p.z_out, *dm = p.depth_map;
            uint8_t *rect = p.rectified;
            if (host) {
                for (int j = 0; j < 5; ++j)
                    if (p.in[j] && p.in_bytes[j])
                        RS_CUDA(ctx, cudaMemcpyAsync(L->stage[j].p, p.in[j], p.in_bytes[j], cudaMemcpyHostToDevice, ctx->s_in));
                RS_CUDA(ctx, cudaEventRecord(L->ev_in[0], ctx->s_in));
                RS_CUDA(ctx, cudaStreamWaitEvent(L->stream, L->ev_in[0], 0));
                for (int j = 0; j < 5; ++j) in[j] = L->stage[j].p;
                z = (double *)L->stage[5].p; dm = (double *)L->stage[6].p; rect = (uint8_t *)L->stage[7].p;
            } else if (!dm) {
                dm = (double *)L->stage[6].p;     // nobody wants the depth map, but the splat reads it
            }
            const StepArgs a = step_args(C, p, in, z, dm, rect);
            const int qrc = queue_step(L, a, p.v, p.w, *p.k);
            if (qrc != RSDSFM_OK) { if (L != ctx) ctx->err = L->err; return qrc; }
            RS_CUDA(ctx, cudaEventRecord(L->ev_cdone[0], L->stream));
            RS_CUDA(ctx, cudaStreamWaitEvent(L->s_out, L->ev_cdone[0], 0));
            if (host) {
                RS_CUDA(ctx, cudaMemcpyAsync(p.z_out, a.z, sizeof(double) * mm, cudaMemcpyDeviceToHost, L->s_out));
                if (p.depth_map)
                    RS_CUDA(ctx, cudaMemcpyAsync(p.depth_map, a.depth_map, sizeof(double) * tot, cudaMemcpyDeviceToHost, L->s_out));
                RS_CUDA(ctx, cudaMemcpyAsync(p.rectified, a.rectified, tot * 3, cudaMemcpyDeviceToHost, L->s_out));
            }
            RS_CUDA(ctx, cudaEventRecord(L->ev_out[0], L->s_out));
            return RSDSFM_OK;
        }();
        if (rc != RSDSFM_OK) {
            drain_all();
            *p.status = rc;
            if (first_err == RSDSFM_OK) first_err = rc;
            continue;
        }
        occupant[l] = i;
    }
    for (int l = 0; l < n_lanes; ++l) {
        const int rc = finish(l);
        if (rc != RSDSFM_OK && first_err == RSDSFM_OK) first_err = rc;
    }
    drain_all();
    if (trace) {
        const auto tc2 = std::chrono::steady_clock::now();
        fprintf(stderr, "[rsdsfm trace] sequence of %d pairs, %d lanes (%d active): setup %.2f ms, pairs %.2f ms\n", n_ok, n_lanes, active,
                std::chrono::duration<double, std::milli>(tc1 - tc0).count(), std::chrono::duration<double, std::milli>(tc2 - tc1).count());
    }
    restore();
    for (int l = 1; l < n_lanes; ++l) {
        rsdsfm_ctx *L = lane[l];
        ctx->launches += L->launches - launches_before[l];
        if (ctx->profile) {                       // the lanes' timers join the caller-visible ones
            for (int j = 0; j < 8; ++j) { ctx->prof[j] += L->prof[j]; ctx->prof_detail[j] += L->prof_detail[j]; L->prof[j] = 0.0; L->prof_detail[j] = 0.0; }
        }
        if (L->exc_cap > ctx->exc_cap) ctx->exc_cap = L->exc_cap;
    }
    return first_err;
}

static SeqPair expanded_view(rsdsfm_pair_io &p, size_t tot)
{
    const size_t mm = (size_t)(p.m > 0 ? p.m : 0);
    SeqPair q{};
    q.in[0] = p.flow; q.in[1] = p.inliers3; q.in[2] = p.alpha; q.in[3] = p.alpha_k; q.in[4] = p.image;
    q.in_bytes[0] = sizeof(double) * 2 * mm; q.in_bytes[1] = sizeof(double) * 3 * mm; q.in_bytes[2] = sizeof(double) * mm;
    q.in_bytes[3] = sizeof(double) * mm; q.in_bytes[4] = tot * 3;
    q.m = p.m; q.n = p.m;
    q.v = p.v; q.w = p.w; q.k = &p.k;
    q.z_out = p.z_out; q.depth_map = p.depth_map; q.rectified = p.rectified;
    q.summary = &p.summary; q.status = &p.status;
    return q;
}

static SeqPair compact_view(rsdsfm_compact_pair_io &p, size_t tot, int flow_f32)
{
    const size_t nn = (size_t)(p.n > 0 ? p.n : 0);
    SeqPair q{};
    q.in[0] = p.flow_img; q.in[1] = p.mask; q.in[2] = p.inv_depth; q.in[3] = nullptr; q.in[4] = p.image;
    q.in_bytes[0] = (flow_f32 ? sizeof(float) : sizeof(double)) * 2 * tot; q.in_bytes[1] = nn; q.in_bytes[2] = sizeof(double) * nn;
    q.in_bytes[3] = 0; q.in_bytes[4] = tot * 3;
    q.m = p.m; q.n = p.n;
    q.v = p.v; q.w = p.w; q.k = &p.k;
    q.z_out = p.z_out; q.depth_map = p.depth_map; q.rectified = p.rectified;
    q.summary = &p.summary; q.status = &p.status;
    return q;
}

// ---- a2..a15 on device pointers ------------------------------------------------------------------
// minimal.cc:226-244: every trial draws 9 indices with rand() % n_temp from an index vector that
// persists (with its swaps) across trials.  The vector is kept sparse: only displaced entries.
static void draws_to_samples(const uint32_t *draws, int H, int n, std::vector<int32_t> &samples)
{
    std::unordered_map<int, int> moved;
    auto at = [&](int i) { auto it = moved.find(i); return it == moved.end() ? i : it->second; };
    samples.resize((size_t)H * 9);
    for (int t = 0; t < H; ++t) {
        int n_temp = n;
        for (int j = 0; j < 9; ++j) {
            const int c = (int)(draws[(size_t)t * 9 + j] % (uint32_t)n_temp);
            const int last = at(n_temp - 1), pick = at(c);
            moved[n_temp - 1] = pick;                                    // std::swap(indices[n_temp-1], indices[c])
            moved[c] = last;
            samples[(size_t)t * 9 + j] = pick;
            n_temp--;
        }
    }
}

__global__ void k_fill_double(double *p, int n, double val)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = val;
}

static int pipeline_device(rsdsfm_ctx *ctx, const rsdsfm_pipeline_params &P, const double *d_flow_img, const uint8_t *d_image,
                           rsdsfm_pipeline_io *io, double *d_dm, uint8_t *d_out)
{
    const size_t tot = (size_t)P.rows * P.cols;
    const int H = P.num_hypotheses;
    io->n = io->m = 0; io->best_idx = -1;
    memset(&io->summary, 0, sizeof io->summary);
    DevBuf *B = ctx->pipe;
    RS_TRY(ensure(ctx, B[0], sizeof(double) * 2 * tot));   // coord
    RS_TRY(ensure(ctx, B[1], sizeof(double) * 2 * tot));   // flow
    RS_TRY(ensure(ctx, B[2], sizeof(double) * 2 * tot));   // coord_px
    RS_TRY(ensure(ctx, B[3], sizeof(double) * 2 * tot));   // flow_px
    RS_TRY(ensure(ctx, B[4], sizeof(int32_t) * tot));      // pixel index
    double *coord = (double *)B[0].p, *flow = (double *)B[1].p, *cpx = (double *)B[2].p, *fpx = (double *)B[3].p;
    int n = 0;
    RS_TRY(flatten_device(ctx, d_flow_img, P.rows, P.cols, P.K4, P.gamma, P.flow_threshold, coord, flow, cpx, fpx,
                          (int32_t *)B[4].p, &n));
    io->n = n;
    // main.cc never truncates (Q3): with emulate_padding the point set is all rows*cols columns, the tail
    // being what k_flat_tail wrote (coord = 1, flow = 0  =>  alpha = 1, alpha_k = the reference's value at y = 1)
    if (P.emulate_padding) n = (int)tot;
    if (n < 9) return fail(ctx, RSDSFM_ERR_ARG, "pipeline: fewer than 9 valid flow vectors");
    const size_t nn = (size_t)n;
    RS_TRY(ensure(ctx, B[5], sizeof(double) * nn));        // alpha
    RS_TRY(ensure(ctx, B[6], sizeof(double) * nn));        // alpha_k
    RS_TRY(ensure(ctx, B[7], nn));                         // consensus mask of the winner
    RS_TRY(ensure(ctx, B[8], sizeof(double) * nn));        // its inverse depths
    RS_TRY(ensure(ctx, B[9], sizeof(double) * 3 * nn));    // inliers3
    RS_TRY(ensure(ctx, B[10], sizeof(double) * nn));       // alpha of the inliers
    RS_TRY(ensure(ctx, B[11], sizeof(double) * nn));       // alpha_k of the inliers
    RS_TRY(ensure(ctx, B[12], sizeof(int32_t) * nn));      // flattened index of the inliers
    RS_TRY(ensure(ctx, B[13], sizeof(double) * nn));       // z
    double *alpha = (double *)B[5].p, *alpha_k = (double *)B[6].p;
    RS_TRY(alpha_device(ctx, fpx, cpx, n, (double)P.rows, P.gamma, alpha, alpha_k));
    if (P.gs_mode) {                                       // alpha *= 0; alpha += 1 (main.cc:441-444)
        k_fill_double<<<grid_for(ctx, n, 4), kThreads, 0, ctx->stream>>>(alpha, n, 1.0);
        ctx->launches++;
    }
    std::vector<int32_t> samples;
    const int32_t *smp = io->samples;
    if (!smp) { draws_to_samples(io->draws, H, n, samples); smp = samples.data(); }
    for (size_t i = 0; i < (size_t)H * 9; ++i)
        if (smp[i] < 0 || smp[i] >= n) return fail(ctx, RSDSFM_ERR_ARG, "pipeline: sample index out of range");
    std::vector<double> hyps((size_t)H * 7);
    std::vector<int> counts((size_t)H);
    std::vector<double> sumerr((size_t)H);
    RS_TRY(ransac_fit_device(ctx, coord, flow, alpha, alpha_k, n, P.const_acceleration, smp, H, hyps.data()));
    int best = -1;
    RS_TRY(ransac_score_device(ctx, coord, flow, alpha, alpha_k, n, hyps.data(), H, P.ransac_tolerance, counts.data(),
                               sumerr.data(), &best, (uint8_t *)B[7].p, (double *)B[8].p));
    io->best_idx = best;
    if (best < 0) return fail(ctx, RSDSFM_ERR_ARG, "pipeline: no RANSAC trial produced a consensus set");
    // calculateVelocities returns (w, v, k); RansacValues stores v, w, k
    double v[3], w[3], k;
    for (int j = 0; j < 3; ++j) { w[j] = hyps[(size_t)best * 7 + j]; v[j] = hyps[(size_t)best * 7 + 3 + j]; }
    k = hyps[(size_t)best * 7 + 6];
    for (int j = 0; j < 3; ++j) { io->ransac_motion[j] = v[j]; io->ransac_motion[3 + j] = w[j]; }
    io->ransac_motion[6] = k;
    int m = 0;
    double *inl = (double *)B[9].p;
    RS_TRY(gather_inliers_device(ctx, coord, alpha, alpha_k, n, (const uint8_t *)B[7].p, (const double *)B[8].p, inl,
                                 (double *)B[10].p, (double *)B[11].p, (int32_t *)B[12].p, &m));
    io->m = m;
    if (m <= 0) return fail(ctx, RSDSFM_ERR_ARG, "pipeline: empty consensus set");
    if (P.use_refinement) {
        StepArgs a{flow, inl, (const double *)B[10].p, (const double *)B[11].p,
                   P.repair_pairing ? (const int32_t *)B[12].p : nullptr, d_image, m, P.const_acceleration, P.gs_mode,
                   P.rows, P.cols, P.layout, P.K4, P.gamma, (double *)B[13].p, d_dm, d_out};
        for (int attempt = 0;; ++attempt) {
            RS_TRY(queue_step(ctx, a, v, w, k));
            RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            bool overflow = false;
            RS_TRY(finish_step(ctx, P.const_acceleration ? 7 : 6, m, v, w, &k, &io->summary, &overflow));
            if (!overflow) break;
            if (attempt == 1) return fail(ctx, RSDSFM_ERR_INTERNAL, "pipeline: exception list overflow persisted");
        }
    } else {
        // results = ransac_results (main.cc:455): row 2 of the inliers is used as it is
        RS_TRY(ensure(ctx, ctx->misc, 256));
        RS_TRY(ensure(ctx, ctx->poses, sizeof(double) * 12 * (size_t)P.rows));
            double *stats = (double *)ctx->misc.p, *dmot = stats + 16;
        double *dR = (double *)ctx->poses.p, *dt = dR + 9 * (size_t)P.rows;
        double *hm = (double *)pinned_lm_init(ctx);
        for (int j = 0; j < 3; ++j) { hm[j] = v[j]; hm[3 + j] = w[j]; }
        hm[6] = k;
        RS_CUDA(ctx, cudaMemcpyAsync(dmot, hm, sizeof(double) * 7, cudaMemcpyHostToDevice, ctx->stream));
        RS_TRY(glue_device(ctx, inl + 2, 3, inl, 3, m, P.K4, P.rows, P.cols, INFINITY, P.layout, d_dm, nullptr, stats));
        RS_TRY(poses_device(ctx, dmot, stats, P.gamma, P.rows, dR, dt));
        RS_TRY(backproject_fill_device(ctx, d_image, d_dm, P.layout, P.rows, P.cols, P.K4, dR, dt, P.gs_mode, d_out));
        RS_CUDA(ctx, cudaMemcpyAsync(pinned_stats(ctx), stats, sizeof(double) * 8, cudaMemcpyDeviceToHost, ctx->stream));
        RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (pinned_stats(ctx)[3] < 0) for (int j = 0; j < 3; ++j) v[j] *= -1.0;
        io->summary.termination = RSDSFM_CONVERGENCE;
    }
    for (int j = 0; j < 3; ++j) { io->v[j] = v[j]; io->w[j] = w[j]; }
    io->k = k;
    return RSDSFM_OK;
}

static bool pipeline_args_ok(const rsdsfm_pipeline_params *P, const rsdsfm_pipeline_io *io)
{
    return P && io && P->rows > 0 && P->cols > 0 && P->num_hypotheses > 0 && io->flow_img && io->image &&
           (io->samples || io->draws) && io->depth_map && io->rectified;
}

}  // namespace rsdsfm

using namespace rsdsfm;

#define RS_ENTER(ctx)                                                     \
    if (!(ctx)) return RSDSFM_ERR_ARG;                                    \
    RS_CUDA((ctx), cudaSetDevice((ctx)->device))

extern "C" {

int rsdsfm_refine_rectify(rsdsfm_ctx *ctx, int mem, const double *flow, const double *inliers3, const double *alpha,
                          const double *alpha_k, int m, double *v, double *w, double *k, int const_acceleration,
                          int gs_mode, const uint8_t *image, int rows, int cols, const double *K4, double gamma, int layout,
                          double *z_out, double *depth_map, uint8_t *rectified, rsdsfm_lm_summary *summary)
{
    RS_ENTER(ctx);
    if (m <= 0 || rows <= 0 || cols <= 0 || !flow || !inliers3 || !alpha || !alpha_k || !v || !w || !k || !image || !K4 ||
        !z_out || !depth_map || !rectified)
        return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_refine_rectify: bad argument");
    rsdsfm_pair_io io{};
    io.flow = flow; io.inliers3 = inliers3; io.alpha = alpha; io.alpha_k = alpha_k; io.image = image; io.m = m;
    io.z_out = z_out; io.depth_map = depth_map; io.rectified = rectified;
    SeqPair p = expanded_view(io, (size_t)rows * cols);
    p.v = v; p.w = w; p.k = k; p.summary = summary;
    const SeqCommon C{0, 0, 0.0, const_acceleration, gs_mode, rows, cols, layout, K4, gamma};
    return step_sync(ctx, mem, 0, C, p);
}

int rsdsfm_refine_rectify_sequence(rsdsfm_ctx *ctx, int mem, int n_pairs, rsdsfm_pair_io *pairs, int const_acceleration,
                                   int gs_mode, int rows, int cols, const double *K4, double gamma, int layout)
{
    RS_ENTER(ctx);
    if (n_pairs < 0 || (n_pairs > 0 && !pairs) || rows <= 0 || cols <= 0 || !K4)
        return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_refine_rectify_sequence: bad argument");
    std::vector<SeqPair> views;
    for (int i = 0; i < n_pairs; ++i) views.push_back(expanded_view(pairs[i], (size_t)rows * cols));
    const SeqCommon C{0, 0, 0.0, const_acceleration, gs_mode, rows, cols, layout, K4, gamma};
    return sequence_core(ctx, mem, views, C);
}

int rsdsfm_refine_rectify_compact_sequence(rsdsfm_ctx *ctx, int mem, int n_pairs, rsdsfm_compact_pair_io *pairs, int flow_f32,
                                           double flow_threshold, int const_acceleration, int gs_mode, int rows, int cols,
                                           const double *K4, double gamma, int layout)
{
    RS_ENTER(ctx);
    if (n_pairs < 0 || (n_pairs > 0 && !pairs) || rows <= 0 || cols <= 0 || !K4)
        return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_refine_rectify_compact_sequence: bad argument");
    std::vector<SeqPair> views;
    for (int i = 0; i < n_pairs; ++i) views.push_back(compact_view(pairs[i], (size_t)rows * cols, flow_f32));
    const SeqCommon C{1, flow_f32 ? 1 : 0, flow_threshold, const_acceleration, gs_mode, rows, cols, layout, K4, gamma};
    if (n_pairs == 1) {                          // a one-pair sequence is the synchronous call
        SeqPair &p = views[0];
        *p.status = seq_pair_ok(C, p) ? RSDSFM_OK : fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_refine_rectify_compact_sequence: bad pair argument");
        if (*p.status == RSDSFM_OK) *p.status = step_sync(ctx, mem, 0, C, p);
        return *p.status;
    }
    return sequence_core(ctx, mem, views, C);
}

int rsdsfm_pipeline_pair(rsdsfm_ctx *ctx, int mem, const rsdsfm_pipeline_params *params, rsdsfm_pipeline_io *io)
{
    RS_ENTER(ctx);
    if (!pipeline_args_ok(params, io)) return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_pipeline_pair: bad argument");
    const size_t tot = (size_t)params->rows * params->cols;
    ctx->io_slot = 0;
    const void *d_fi = nullptr, *d_img = nullptr;
    void *d_dm = nullptr, *d_out = nullptr;
    RS_TRY(stage_in(ctx, mem, 0, io->flow_img, sizeof(double) * 2 * tot, &d_fi));
    RS_TRY(stage_in(ctx, mem, 4, io->image, tot * 3, &d_img));
    RS_TRY(stage_out_reserve(ctx, mem, 6, io->depth_map, sizeof(double) * tot, &d_dm));
    RS_TRY(stage_out_reserve(ctx, mem, 7, io->rectified, tot * 3, &d_out));
    io->status = pipeline_device(ctx, *params, (const double *)d_fi, (const uint8_t *)d_img, io, (double *)d_dm, (uint8_t *)d_out);
    if (io->status != RSDSFM_OK) { cudaStreamSynchronize(ctx->stream); return io->status; }
    RS_TRY(stage_out(ctx, mem, io->depth_map, d_dm, sizeof(double) * tot));
    RS_TRY(stage_out(ctx, mem, io->rectified, d_out, tot * 3));
    if (mem == RSDSFM_HOST) RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RSDSFM_OK;
}

int rsdsfm_pipeline_sequence(rsdsfm_ctx *ctx, int mem, const rsdsfm_pipeline_params *params, int n_pairs,
                             rsdsfm_pipeline_io *pairs)
{
    RS_ENTER(ctx);
    if (!params || n_pairs < 0 || (n_pairs > 0 && !pairs)) return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_pipeline_sequence: bad argument");
    int first_err = RSDSFM_OK;
    if (mem != RSDSFM_HOST) {                    // nothing to overlap: the stages synchronise for their counts anyway
        for (int i = 0; i < n_pairs; ++i) {
            int rc = rsdsfm_pipeline_pair(ctx, mem, params, &pairs[i]);
            pairs[i].status = rc;
            if (rc != RSDSFM_OK && first_err == RSDSFM_OK) first_err = rc;
        }
        return first_err;
    }
    RS_TRY(ensure_io(ctx));
    const size_t tot = (size_t)params->rows * params->cols;
    drain(ctx);
    for (int s = 0; s < 2; ++s) {
        RS_TRY(ensure(ctx, ctx->stage[8 * s + 0], sizeof(double) * 2 * tot));
        RS_TRY(ensure(ctx, ctx->stage[8 * s + 4], tot * 3));
        RS_TRY(ensure(ctx, ctx->stage[8 * s + 6], sizeof(double) * tot));
        RS_TRY(ensure(ctx, ctx->stage[8 * s + 7], tot * 3));
    }
    // valid pairs, in order; position q in this list occupies I/O slot q&1
    std::vector<int> idx;
    for (int i = 0; i < n_pairs; ++i) {
        if (pipeline_args_ok(params, &pairs[i])) { idx.push_back(i); continue; }
        pairs[i].status = fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_pipeline_sequence: bad pair argument");
        if (first_err == RSDSFM_OK) first_err = RSDSFM_ERR_ARG;
    }
    auto upload = [&](size_t q) -> int {         // flow field and frame of list entry q -> its slot, on the upload stream
        const int s = (int)(q & 1), b = 8 * s;
        const rsdsfm_pipeline_io &p = pairs[idx[q]];
        RS_CUDA(ctx, cudaMemcpyAsync(ctx->stage[b + 0].p, p.flow_img, sizeof(double) * 2 * tot, cudaMemcpyHostToDevice, ctx->s_in));
        RS_CUDA(ctx, cudaMemcpyAsync(ctx->stage[b + 4].p, p.image, tot * 3, cudaMemcpyHostToDevice, ctx->s_in));
        RS_CUDA(ctx, cudaEventRecord(ctx->ev_in[s], ctx->s_in));
        return RSDSFM_OK;
    };
    bool out_pending[2] = {false, false};
    if (!idx.empty()) RS_TRY(upload(0));
    for (size_t q = 0; q < idx.size(); ++q) {
        rsdsfm_pipeline_io &p = pairs[idx[q]];
        const int s = (int)(q & 1), b = 8 * s;
        // the other slot's previous occupant (entry q-1) has finished computing -- every pair ends
        // synchronised -- so the next entry's inputs can go up while this one computes
        if (q + 1 < idx.size()) RS_TRY(upload(q + 1));
        RS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[s], 0));
        if (out_pending[s]) RS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_out[s], 0));   // entry q-2's download
        ctx->io_slot = s;
        const int rc = pipeline_device(ctx, *params, (const double *)ctx->stage[b + 0].p, (const uint8_t *)ctx->stage[b + 4].p, &p,
                                       (double *)ctx->stage[b + 6].p, (uint8_t *)ctx->stage[b + 7].p);
        p.status = rc;
        if (rc != RSDSFM_OK) {
            cudaStreamSynchronize(ctx->stream);
            if (first_err == RSDSFM_OK) first_err = rc;
            continue;
        }
        // compute of this entry has completed: download on s_out while the next entry computes
        RS_CUDA(ctx, cudaMemcpyAsync(p.depth_map, ctx->stage[b + 6].p, sizeof(double) * tot, cudaMemcpyDeviceToHost, ctx->s_out));
        RS_CUDA(ctx, cudaMemcpyAsync(p.rectified, ctx->stage[b + 7].p, tot * 3, cudaMemcpyDeviceToHost, ctx->s_out));
        RS_CUDA(ctx, cudaEventRecord(ctx->ev_out[s], ctx->s_out));
        out_pending[s] = true;
    }
    drain(ctx);
    ctx->io_slot = 0;
    return first_err;
}

}  // extern "C"
