// refine.cu -- GPU Levenberg-Marquardt for the reference's two Ceres problems:
//   a8  nonlinear_refinement::estimateInverseDepths  (nonlinearRefinement.cc:109-180)  NF = 0
//   a9  nonlinear_refinement::nonLinearRefinement    (nonlinearRefinement.cc:183-252)  NF = 6 | 7
//
// ONE persistent cooperative kernel (k_lm_solve, lm_kernel.cuh) runs the whole solve; this file holds
// the kernels that lay the inputs out for it and the host side (launch, control block, read-back).
//
// Data layout in HBM: tile-blocked structure of arrays, tile = kTile residual blocks:
//   blk[tile] = { xy[kTile] (x, y) | uu[kTile] (ux, uy; Q1 pairing applied once by the gather kernel)
//                 | aa[kTile] (alpha, alpha_k) } as double2, contiguous = ONE bulk copy per tile,
//   d[2][tiles*kTile] inverse depth ping-pong (x / candidate).
// The last tile is padded with all-zero records (k_lm_pad): they contribute nothing to any sum.
// Traffic per residual block: INIT reads 56 B, FUSED reads 56 B and writes 8 B (algorithmic
// minimum, SURVEY.md 8d: 24 B / 56 B -- x, y, alpha, alpha_k are inputs of the C ABI).
#include "lm_kernel.cuh"

namespace rsdsfm {

// ------------------------------------------------------------------------------------------
// gather: API arrays -> SoA records.  Residual i pairs inlier i with flow(:, i) of the array the
// caller passed (reference behaviour, nonlinearRefinement.cc:209-212) or flow(:, flow_index[i]).
// ------------------------------------------------------------------------------------------
__global__ void k_refine_gather(const double *__restrict__ flow, const double *__restrict__ inliers3,
                                const double *__restrict__ alpha, const double *__restrict__ alpha_k,
                                const int32_t *__restrict__ flow_index, int m, double2 *blk, double *d0, LmShared *sh)
{
    int bad = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const double x = inliers3[3 * (size_t)i], y = inliers3[3 * (size_t)i + 1], z = inliers3[3 * (size_t)i + 2];
        const int fi = flow_index ? flow_index[i] : i;
        const double2 u = reinterpret_cast<const double2 *>(flow)[fi];
        blk[blk_index(i, 0)] = make_double2(x, y);
        blk[blk_index(i, 1)] = u;
        blk[blk_index(i, 2)] = make_double2(alpha[i], alpha_k[i]);
        const double d = 1.0 / z;                              // nonlinearRefinement.cc:213
        d0[i] = d;
        if (!isfinite(d)) bad = 1;
    }
    // solver.cc: non-finite initial parameter values => FAILURE before any evaluation
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(&sh->nonfinite_input, 1);
}

// a8 variant: depth starts at 1.0 (:140); coord / flow are copied so the solver's 16-byte
// alignment requirement never leaks into the C ABI.
__global__ void k_depth_gather(const double *__restrict__ coord, const double *__restrict__ flow,
                               const double *__restrict__ alpha, const double *__restrict__ alpha_k, int n, double2 *blk,
                               double *d0)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        blk[blk_index(i, 0)] = make_double2(coord[2 * (size_t)i], coord[2 * (size_t)i + 1]);
        blk[blk_index(i, 1)] = make_double2(flow[2 * (size_t)i], flow[2 * (size_t)i + 1]);
        blk[blk_index(i, 2)] = make_double2(alpha[i], alpha_k[i]);
        d0[i] = 1.0;
    }
}

// all-zero records behind the last residual block of the last tile, in blk and in both depth planes
__global__ void k_lm_pad(double2 *blk, double *d0, double *d1, int m)
{
    const int end = ((m + kTile - 1) / kTile) * kTile;
    for (int i = m + (int)threadIdx.x; i < end; i += (int)blockDim.x) {
        blk[blk_index(i, 0)] = make_double2(0.0, 0.0);
        blk[blk_index(i, 1)] = make_double2(0.0, 0.0);
        blk[blk_index(i, 2)] = make_double2(0.0, 0.0);
        d0[i] = 0.0; d1[i] = 0.0;
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int NF>
static int launch_persistent(rsdsfm_ctx *ctx, RefineData D, double *d0, double *d1, LmShared *sh, double *partials,
                             ExcEntry *exc, unsigned int exc_cap, const double *z_in, int z_stride, double *out,
                             int invert_out, double *zstats, int grid)
{
    const size_t smem = sizeof(Stage) * (size_t)kStages;
    if (ctx->profile) cudaEventRecord(ctx->pe0[ctx->io_slot], ctx->stream);
    RS_CUDA(ctx, cudaFuncSetAttribute(k_lm_solve<NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SolveArgs a{};
    a.D = D; a.d0 = d0; a.d1 = d1; a.sh = sh; a.partials = partials; a.exc = exc; a.exc_cap = exc_cap;
    a.z_in = z_in; a.z_stride = z_stride; a.out = out; a.invert_out = invert_out; a.zstats = zstats;
    a.peers.n = 1; a.peers.me = 0; a.peers.epoch = 0;
    if (ctx->n_peers > 1) {            // row split: every GPU of the group launches the same solve (collective call)
        a.peers.n = ctx->n_peers; a.peers.me = ctx->my_peer; a.peers.epoch = ++ctx->peer_epoch;
        for (int g = 0; g < ctx->n_peers; ++g) a.peers.mail[g] = (Tagged *)ctx->peer_mail[g];
    }
    void *args[] = {&a};
    RS_CUDA(ctx, cudaLaunchCooperativeKernel((void *)k_lm_solve<NF>, dim3(grid), dim3(kThreads), args, smem, ctx->stream));
    if (ctx->profile) cudaEventRecord(ctx->pe1[ctx->io_slot], ctx->stream);
    ctx->launches++;
    return RSDSFM_OK;
}

// The initial control block travels as a kernel parameter and the final one is written straight
// into pinned host memory by a kernel: neither transfer queues behind the large uploads /
// downloads that a pipelined sequence keeps on the copy engines.
static_assert(sizeof(LmShared) <= 3840 && sizeof(LmShared) % 4 == 0, "control block must fit the kernel parameter space");

__global__ void __launch_bounds__(256) k_lm_begin(LmShared *sh, const __grid_constant__ LmShared init, int keep_input_flag)
{
    // nonfinite_input is the LAST field: with keep_input_flag the gather kernel's verdict survives
    const int words = (int)((keep_input_flag ? offsetof(LmShared, nonfinite_input) : sizeof(LmShared)) / 4);
    const unsigned int *src = reinterpret_cast<const unsigned int *>(&init);
    unsigned int *dst = reinterpret_cast<unsigned int *>(sh);
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    // "gather found a non-finite start depth" => the FAILURE Ceres reports before evaluating anything
    if (threadIdx.x == 0 && keep_input_flag && sh->nonfinite_input) {
        sh->bc.next = LM_DONE;
        sh->ctl.termination = RSDSFM_FAILURE;
        sh->ctl.reason = RSDSFM_REASON_NONFINITE_INPUT;
    }
}

__global__ void __launch_bounds__(256) k_lm_readback(const LmShared *sh, const double *stats8, LmShared *host_block,
                                                     double *host_stats8)
{
    const unsigned int *src = reinterpret_cast<const unsigned int *>(sh);
    unsigned int *dst = reinterpret_cast<unsigned int *>(host_block);            // mapped pinned memory (UVA)
    for (int i = threadIdx.x; i < (int)(sizeof(LmShared) / 4); i += blockDim.x) dst[i] = src[i];
    if (stats8 && threadIdx.x < 8) host_stats8[threadIdx.x] = stats8[threadIdx.x];
}

int lm_grid_size(const rsdsfm_ctx *ctx)
{
    static const int forced = getenv("RSDSFM_LM_GRID") ? atoi(getenv("RSDSFM_LM_GRID")) : 0;      // experiments (tools/lm_bench.py)
    const int g = forced > 0 ? forced : ctx->lm_grid;
    return (g > 0 && g < ctx->num_sms) ? g : ctx->num_sms;
}

// Initial capacity (entries) of each clamped-pixel list; an overflow is detected by the kernel and the
// solve is repeated with room for every residual block.  RSDSFM_EXC_CAP overrides the default so that
// the tests can drive the overflow path with a handful of clamped pixels.
static int min_exc_cap()
{
    const char *e = getenv("RSDSFM_EXC_CAP");
    const int v = e ? atoi(e) : 0;
    return v > 0 ? v : 4096;
}

// Queues one LM solve on the context's stream (no host synchronisation).  The control block
// (ctx->lm_shared) receives the result; lm_collect() reads it back.
static int lm_solve_async(rsdsfm_ctx *ctx, const RefineData &D, double *d0, double *d1, int nf, const Motion &mot0,
                          const rsdsfm_lm_options &opt, const double *z_in, int z_stride, double *out, int invert_out,
                          bool keep_input_flag, double *zstats = nullptr)
{
    RS_TRY(ensure(ctx, ctx->lm_shared, sizeof(LmShared)));
    LmShared *sh = (LmShared *)ctx->lm_shared.p;
    const int grid = lm_grid_size(ctx);                 // one persistent CTA per SM (a fraction of them on a lane of a sequence)
    const int nv = (nf == 0) ? Row<0>::NV : (nf == 6 ? Row<6>::NV : Row<7>::NV);
    RS_TRY(ensure(ctx, ctx->partials, sizeof(double) * 2 * (size_t)kStrips * nv));   // rows (one per strip) of even / odd phases
    if (ctx->exc_cap < min_exc_cap()) ctx->exc_cap = min_exc_cap();
    // (row split: a member cannot repeat its solve alone when its list overflows, so the lists start larger)
    if (ctx->n_peers > 1 && ctx->exc_cap < D.m / 16 + 4096) ctx->exc_cap = D.m / 16 + 4096;
    RS_TRY(ensure(ctx, ctx->exc, sizeof(ExcEntry) * kExcSlots * (size_t)ctx->exc_cap));   // current + speculative + being cleared
    static_assert(sizeof(LmShared) <= 8192 - 256, "pinned slot layout (common.cuh)");

    // initial control block: built on the host, passed by value to k_lm_begin
    LmShared init_block;
    LmShared *h = &init_block;
    memset(h, 0, sizeof(LmShared));
    double f0[kMaxNF] = {0, 0, 0, 0, 0, 0, 0};
    if (nf >= 6) { for (int j = 0; j < 3; ++j) { f0[j] = mot0.v[j]; f0[3 + j] = mot0.w[j]; } }
    if (nf == 7) f0[6] = mot0.k;
    h->ctl.init(opt, nf, f0);
    h->base = mot0; h->bc.mot = mot0; h->bc.cand = mot0;
    h->bc.next = LM_RUN_A; h->bc.which_x = 0; h->bc.first = 1;
    h->bc.radius = h->ctl.radius;
    bool finite = isfinite(mot0.k);
    for (int j = 0; j < 3; ++j) finite = finite && isfinite(mot0.v[j]) && isfinite(mot0.w[j]);
    if (!finite) {              // solver.cc: non-finite parameter blocks => FAILURE, nothing evaluated
        h->bc.next = LM_DONE; h->ctl.termination = RSDSFM_FAILURE; h->ctl.reason = RSDSFM_REASON_NONFINITE_INPUT;
    }
    k_lm_begin<<<1, 256, 0, ctx->stream>>>(sh, *h, keep_input_flag ? 1 : 0);
    k_lm_pad<<<1, 256, 0, ctx->stream>>>(const_cast<double2 *>(D.blk), d0, d1, D.m);
    ctx->launches += 2;
    double *partials = (double *)ctx->partials.p;
    ExcEntry *exc = (ExcEntry *)ctx->exc.p;
    const unsigned int cap = (unsigned int)ctx->exc_cap;
    if (nf == 0) return launch_persistent<0>(ctx, D, d0, d1, sh, partials, exc, cap, z_in, z_stride, out, invert_out, zstats, grid);
    if (nf == 6) return launch_persistent<6>(ctx, D, d0, d1, sh, partials, exc, cap, z_in, z_stride, out, invert_out, zstats, grid);
    return launch_persistent<7>(ctx, D, d0, d1, sh, partials, exc, cap, z_in, z_stride, out, invert_out, zstats, grid);
}

// Queues the read-back of the control block (and, if given, of the 8 depth statistics) into the
// current I/O slot's pinned area.
int lm_collect_enqueue(rsdsfm_ctx *ctx, const double *stats_dev)
{
    k_lm_readback<<<1, 256, 0, ctx->stream>>>((const LmShared *)ctx->lm_shared.p, stats_dev, (LmShared *)pinned_lm_result(ctx),
                                              pinned_stats(ctx));
    ctx->launches++;
    RS_CUDA(ctx, cudaGetLastError());
    return RSDSFM_OK;
}

// Parses the control block of the current I/O slot once its read-back has completed.
// RSDSFM_ERR_INTERNAL when the in-kernel watchdog tripped; *overflow: the exception list was too
// small (ctx->exc_cap was raised: run the solve again).
int lm_collect_finish(rsdsfm_ctx *ctx, int nf, int m, Motion *mot, rsdsfm_lm_summary *summary, bool *overflow)
{
    LmShared *h = (LmShared *)pinned_lm_result(ctx);
    if (h->error) return fail(ctx, RSDSFM_ERR_INTERNAL, "LM kernel: grid barrier watchdog tripped");
    if (h->nonfinite_input & 2)
        return fail(ctx, RSDSFM_ERR_ARG, "compact input: n / m do not match the flow field and the consensus mask");
    *overflow = h->exc_overflow != 0;
    if (*overflow) { if (ctx->exc_cap < m + 1024) ctx->exc_cap = m + 1024; return RSDSFM_OK; }
    if (getenv("RSDSFM_TRACE"))
        fprintf(stderr, "[rsdsfm trace] ctx %p solve m=%d it=%d t0=%llu t1=%llu (%.1f us; before the first phase %.1f, after the last %.1f)\n", (void *)ctx, m,
                h->ctl.iteration, h->t_abs[0], h->t_abs[1], (double)(h->t_abs[1] - h->t_abs[0]) * 1e-3, (double)(h->t_abs[2] - h->t_abs[0]) * 1e-3,
                (double)(h->t_abs[1] - h->t_abs[3]) * 1e-3);
    float kms = 0.f;
    if (ctx->profile) cudaEventElapsedTime(&kms, ctx->pe0[ctx->io_slot], ctx->pe1[ctx->io_slot]);
    if (summary) {
        h->ctl.fill_summary(summary);
        summary->device_ms = (double)(h->t_phase[0] + h->t_phase[2]) * 1e-6;
    }
    if (ctx->profile) {
        ctx->prof[0] += (double)h->t_phase[0] * 1e-6; ctx->prof[1] += (double)h->t_phase[1];
        ctx->prof[2] += (double)h->t_phase[1] * m;
        ctx->prof[3] += (double)h->t_phase[2] * 1e-6; ctx->prof[4] += (double)h->t_phase[3];
        ctx->prof[5] += (double)h->t_phase[3] * m;
        ctx->prof[6] += kms; ctx->prof[7] += 1.0;
        for (int j = 0; j < 8; ++j) ctx->prof_detail[j] += (double)h->t_phase[4 + j] * 1e-6;
    }
    if (mot && h->ctl.termination != RSDSFM_FAILURE) {
        if (nf >= 6) for (int j = 0; j < 3; ++j) { mot->v[j] = h->ctl.f[j]; mot->w[j] = h->ctl.f[3 + j]; }
        if (nf == 7) mot->k = h->ctl.f[6];
    }
    return RSDSFM_OK;
}

// Reads the control block back (synchronises the stream).
int lm_collect(rsdsfm_ctx *ctx, int nf, int m, Motion *mot, rsdsfm_lm_summary *summary, bool *overflow)
{
    RS_TRY(lm_collect_enqueue(ctx, nullptr));
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return lm_collect_finish(ctx, nf, m, mot, summary, overflow);
}

// Device address of the refined motion (v[3], w[3], k) inside the control block, for the
// rectification stage that follows on the same stream.
const double *lm_motion_device(rsdsfm_ctx *ctx)
{
    return reinterpret_cast<const double *>((const char *)ctx->lm_shared.p + offsetof(LmShared, bc) + offsetof(Bcast, mot));
}

static int ensure_lm_buffers(rsdsfm_ctx *ctx, size_t mm)
{
    const size_t tiles = (mm + kTile - 1) / kTile;             // everything is padded to whole tiles
    RS_TRY(ensure(ctx, ctx->pix, sizeof(double2) * 3 * kTile * tiles));
    RS_TRY(ensure(ctx, ctx->dA, sizeof(double) * kTile * tiles));
    RS_TRY(ensure(ctx, ctx->dB, sizeof(double) * kTile * tiles));
    RS_TRY(ensure(ctx, ctx->lm_shared, sizeof(LmShared)));
    return RSDSFM_OK;
}

// Pre-sizes every solver buffer for up to m residual blocks so that a pipelined sequence does not
// reallocate (and therefore synchronise) between pairs.
int lm_reserve(rsdsfm_ctx *ctx, int m)
{
    RS_TRY(ensure_lm_buffers(ctx, (size_t)(m > 0 ? m : 1)));
    RS_TRY(ensure(ctx, ctx->partials, sizeof(double) * 2 * (size_t)kStrips * Row<7>::NV));
    if (ctx->exc_cap < min_exc_cap()) ctx->exc_cap = min_exc_cap();
    return ensure(ctx, ctx->exc, sizeof(ExcEntry) * kExcSlots * (size_t)ctx->exc_cap);
}

// a9 on device pointers: queues gather + solve on the stream, no synchronisation.
int refine_async(rsdsfm_ctx *ctx, const double *flow, const double *inliers3, const double *alpha,
                 const double *alpha_k, int m, const double *v, const double *w, double k, int const_acc,
                 const int32_t *flow_index, const rsdsfm_lm_options *opts, double *z_out, double *zstats)
{
    rsdsfm_lm_options o;
    if (opts) o = *opts; else rsdsfm_lm_default_options(&o);
    const size_t mm = (size_t)m;
    RS_TRY(ensure_lm_buffers(ctx, mm));
    double2 *blk = (double2 *)ctx->pix.p;
    double *d0 = (double *)ctx->dA.p, *d1 = (double *)ctx->dB.p;
    LmShared *sh = (LmShared *)ctx->lm_shared.p;
    RS_CUDA(ctx, cudaMemsetAsync(&sh->nonfinite_input, 0, sizeof(int), ctx->stream));
    k_refine_gather<<<grid_for(ctx, m, 8), kThreads, 0, ctx->stream>>>(flow, inliers3, alpha, alpha_k, flow_index, m, blk, d0, sh);
    ctx->launches++;
    RefineData D{blk, m};
    Motion mot;
    for (int j = 0; j < 3; ++j) { mot.v[j] = v[j]; mot.w[j] = w[j]; }
    mot.k = k;
    return lm_solve_async(ctx, D, d0, d1, const_acc ? 7 : 6, mot, o, inliers3 + 2, 3, z_out, 1, true, zstats);
}

// The solver's input buffers for m residual blocks, for a producer that fills them itself
// (preproc.cu: compact_build_device): tile-blocked blk, the start inverse depths d0, and the
// device flag the producer raises for bad input (1: non-finite start depth, 2: count mismatch).
int lm_input_buffers(rsdsfm_ctx *ctx, int m, void **blk, double **d0, int **input_flag)
{
    RS_TRY(ensure_lm_buffers(ctx, (size_t)(m > 0 ? m : 1)));
    LmShared *sh = (LmShared *)ctx->lm_shared.p;
    RS_CUDA(ctx, cudaMemsetAsync(&sh->nonfinite_input, 0, sizeof(int), ctx->stream));
    *blk = ctx->pix.p; *d0 = (double *)ctx->dA.p; *input_flag = &sh->nonfinite_input;
    return RSDSFM_OK;
}

// a9 on buffers filled through lm_input_buffers: queues the solve, no synchronisation.
// z_in[m]: the start depths (restored on FAILURE, like Ceres restores its parameter blocks).
int refine_prepared_async(rsdsfm_ctx *ctx, int m, const double *v, const double *w, double k, int const_acc,
                          const rsdsfm_lm_options *opts, const double *z_in, double *z_out, double *zstats)
{
    rsdsfm_lm_options o;
    if (opts) o = *opts; else rsdsfm_lm_default_options(&o);
    RefineData D{(const double2 *)ctx->pix.p, m};
    Motion mot;
    for (int j = 0; j < 3; ++j) { mot.v[j] = v[j]; mot.w[j] = w[j]; }
    mot.k = k;
    return lm_solve_async(ctx, D, (double *)ctx->dA.p, (double *)ctx->dB.p, const_acc ? 7 : 6, mot, o, z_in, 1, z_out, 1, true, zstats);
}

// a9, synchronous: returns the refined motion and the summary on the host.
int refine_device(rsdsfm_ctx *ctx, const double *flow, const double *inliers3, const double *alpha,
                  const double *alpha_k, int m, double *v, double *w, double *k, int const_acc,
                  const int32_t *flow_index, const rsdsfm_lm_options *opts, double *z_out, rsdsfm_lm_summary *summary)
{
    rsdsfm_lm_summary local;
    if (!summary) summary = &local;
    memset(summary, 0, sizeof *summary);
    if (m == 0) { summary->termination = RSDSFM_CONVERGENCE; summary->reason = RSDSFM_REASON_FUNCTION_TOL; return RSDSFM_OK; }
    for (int attempt = 0; attempt < 2; ++attempt) {
        RS_TRY(refine_async(ctx, flow, inliers3, alpha, alpha_k, m, v, w, *k, const_acc, flow_index, opts, z_out, nullptr));
        Motion mot;
        for (int j = 0; j < 3; ++j) { mot.v[j] = v[j]; mot.w[j] = w[j]; }
        mot.k = *k;
        bool overflow = false;
        RS_TRY(lm_collect(ctx, const_acc ? 7 : 6, m, &mot, summary, &overflow));
        if (overflow && ctx->n_peers > 1)
            return fail(ctx, RSDSFM_ERR_INTERNAL, "refine (row split): clamped-pixel list overflow on this member");
        if (overflow) continue;                         // exception list enlarged: run again
        for (int j = 0; j < 3; ++j) { v[j] = mot.v[j]; w[j] = mot.w[j]; }
        *k = mot.k;
        return RSDSFM_OK;
    }
    return fail(ctx, RSDSFM_ERR_INTERNAL, "refine: exception list overflow persisted");
}

// a8 on device pointers.
int estimate_inverse_depths_device(rsdsfm_ctx *ctx, const double *coord, const double *flow, int n, const double *v,
                                   const double *w, double k, const double *alpha, const double *alpha_k,
                                   double *inv_depth, rsdsfm_lm_summary *summary)
{
    rsdsfm_lm_options o;
    rsdsfm_lm_default_options(&o);
    rsdsfm_lm_summary local;
    if (!summary) summary = &local;
    memset(summary, 0, sizeof *summary);
    if (n == 0) { summary->termination = RSDSFM_CONVERGENCE; summary->reason = RSDSFM_REASON_FUNCTION_TOL; return RSDSFM_OK; }
    const size_t nn = (size_t)n;
    RS_TRY(ensure_lm_buffers(ctx, nn));
    double2 *blk = (double2 *)ctx->pix.p;
    double *d0 = (double *)ctx->dA.p, *d1 = (double *)ctx->dB.p;
    k_depth_gather<<<grid_for(ctx, n, 8), kThreads, 0, ctx->stream>>>(coord, flow, alpha, alpha_k, n, blk, d0);
    ctx->launches++;
    RefineData D{blk, n};
    Motion mot;
    for (int j = 0; j < 3; ++j) { mot.v[j] = v[j]; mot.w[j] = w[j]; }
    mot.k = k;
    RS_TRY(lm_solve_async(ctx, D, d0, d1, 0, mot, o, nullptr, 1, inv_depth, 0, false));
    bool overflow = false;
    return lm_collect(ctx, 0, n, nullptr, summary, &overflow);
}

// ---- row split over GPUs: the mailbox of this context and the mailboxes of its peers
size_t lm_mailbox_bytes() { return sizeof(Tagged) * (size_t)kMailSlots * kMaxPeers * kMailLd; }
static_assert(kMailLd >= kRowLd, "a mailbox row holds a row of sums");

}  // namespace rsdsfm
