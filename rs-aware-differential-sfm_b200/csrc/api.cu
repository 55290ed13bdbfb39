// api.cu -- the extern "C" layer of include/rsdsfm.h: context management and the per-stage entry
// points with host<->device staging (the fused drivers live in pipeline.cu).  No CPU fallback:
// every compute entry point needs a usable CUDA device and fails with RSDSFM_ERR_CUDA otherwise.
#include "stages.h"
#include "solve9.h"

namespace rsdsfm {

thread_local std::string g_create_error;

static int finish_host_call(rsdsfm_ctx *ctx, int mem)
{
    if (mem == RSDSFM_HOST) RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RSDSFM_OK;
}

}  // namespace rsdsfm

using namespace rsdsfm;

#define RS_ENTER(ctx)                                                     \
    if (!(ctx)) return RSDSFM_ERR_ARG;                                    \
    RS_CUDA((ctx), cudaSetDevice((ctx)->device))

extern "C" {

int rsdsfm_version(void) { return RSDSFM_VERSION; }

void rsdsfm_lm_default_options(rsdsfm_lm_options *o)
{
    o->max_num_iterations = 50;
    o->function_tolerance = 1e-6;
    o->gradient_tolerance = 1e-10;
    o->parameter_tolerance = 1e-8;
    o->initial_trust_region_radius = 1e4;
    o->max_trust_region_radius = 1e16;
    o->min_trust_region_radius = 1e-32;
    o->min_relative_decrease = 1e-3;
    o->min_lm_diagonal = 1e-6;
    o->max_lm_diagonal = 1e32;
    o->max_num_consecutive_invalid_steps = 5;
}

int rsdsfm_create(int device, void *cuda_stream, rsdsfm_ctx **out)
{
    if (!out) return RSDSFM_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return fail(nullptr, RSDSFM_ERR_CUDA, "rsdsfm_create: no usable CUDA device (this library has no CPU fallback)", e);
    if (device < 0 || device >= count) return fail(nullptr, RSDSFM_ERR_ARG, "rsdsfm_create: device ordinal out of range");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, RSDSFM_ERR_CUDA, "cudaSetDevice", e);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, RSDSFM_ERR_CUDA, "cudaGetDeviceProperties", e);
    if (prop.major != 10)
        return fail(nullptr, RSDSFM_ERR_CUDA, "rsdsfm_create: kernels are built for sm_100a (B200) only");
    rsdsfm_ctx *ctx = new rsdsfm_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    if (cuda_stream) { ctx->stream = (cudaStream_t)cuda_stream; ctx->owns_stream = false; }
    else {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete ctx; return fail(nullptr, RSDSFM_ERR_CUDA, "cudaStreamCreate", e); }
        ctx->owns_stream = true;
    }
    for (int j = 0; j < 2; ++j) { cudaEventCreate(&ctx->pe0[j]); cudaEventCreate(&ctx->pe1[j]); }
    e = cudaMallocHost(&ctx->pinned_io, 2 * kPinnedSlotBytes);
    if (e != cudaSuccess) { rsdsfm_destroy(ctx); return fail(nullptr, RSDSFM_ERR_NOMEM, "cudaMallocHost", e); }
    memset(ctx->pinned_io, 0, 2 * kPinnedSlotBytes);
    *out = ctx;
    return RSDSFM_OK;
}

void rsdsfm_destroy(rsdsfm_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &l : ctx->lanes) if (l) { rsdsfm_destroy(l); l = nullptr; }
    if (ctx->s_in) { cudaStreamSynchronize(ctx->s_in); cudaStreamDestroy(ctx->s_in); }
    if (ctx->s_out) { cudaStreamSynchronize(ctx->s_out); cudaStreamDestroy(ctx->s_out); }
    for (int j = 0; j < 2; ++j) {
        if (ctx->ev_in[j]) cudaEventDestroy(ctx->ev_in[j]);
        if (ctx->ev_cdone[j]) cudaEventDestroy(ctx->ev_cdone[j]);
        if (ctx->ev_out[j]) cudaEventDestroy(ctx->ev_out[j]);
        if (ctx->pe0[j]) cudaEventDestroy(ctx->pe0[j]);
        if (ctx->pe1[j]) cudaEventDestroy(ctx->pe1[j]);
    }
    rsdsfm_peer_disconnect(ctx);
    if (ctx->mailbox) cudaFree(ctx->mailbox);
    if (ctx->pinned_io) cudaFreeHost(ctx->pinned_io);
    DevBuf *all[] = {&ctx->partials, &ctx->sums, &ctx->pix, &ctx->dA, &ctx->dB, &ctx->rdepth, &ctx->misc, &ctx->winner,
                     &ctx->poses, &ctx->hyp, &ctx->rpart, &ctx->scan,
                     &ctx->lm_shared, &ctx->exc, &ctx->splat_tab, &ctx->flow_t};
    for (DevBuf *b : all) if (b->p) cudaFree(b->p);
    for (auto &b : ctx->stage) if (b.p) cudaFree(b.p);
    for (auto &b : ctx->pipe) if (b.p) cudaFree(b.p);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// ---- row split of one solve over several GPUs: mailboxes in peer memory (lm_kernel.cuh: peer_allreduce)
static int peer_mailbox(rsdsfm_ctx *ctx)
{
    RS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->mailbox) {
        cudaError_t e = cudaMalloc(&ctx->mailbox, lm_mailbox_bytes());      // its own allocation: exported through CUDA IPC
        if (e != cudaSuccess) return fail(ctx, RSDSFM_ERR_NOMEM, "cudaMalloc (mailbox)", e);
    }
    RS_CUDA(ctx, cudaMemsetAsync(ctx->mailbox, 0, lm_mailbox_bytes(), ctx->stream));
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RSDSFM_OK;
}

int rsdsfm_peer_disconnect(rsdsfm_ctx *ctx)
{
    if (!ctx) return RSDSFM_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int g = 0; g < 8; ++g) {
        if (ctx->peer_ipc[g] && ctx->peer_mail[g]) cudaIpcCloseMemHandle(ctx->peer_mail[g]);
        ctx->peer_mail[g] = nullptr; ctx->peer_ipc[g] = false;
    }
    ctx->n_peers = 1; ctx->my_peer = 0; ctx->peer_epoch = 0;
    return RSDSFM_OK;
}

int rsdsfm_peer_export(rsdsfm_ctx *ctx, void *handle_out)
{
    RS_ENTER(ctx);
    if (!handle_out) return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_peer_export: null handle");
    static_assert(sizeof(cudaIpcMemHandle_t) == RSDSFM_PEER_HANDLE_BYTES, "IPC handle size");
    RS_TRY(peer_mailbox(ctx));
    cudaIpcMemHandle_t h;
    RS_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->mailbox));
    memcpy(handle_out, &h, sizeof h);
    return RSDSFM_OK;
}

int rsdsfm_peer_connect(rsdsfm_ctx *ctx, int n_members, int my_index, const void *handles)
{
    RS_ENTER(ctx);
    if (n_members < 1 || n_members > 8 || my_index < 0 || my_index >= n_members || !handles)
        return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_peer_connect: 1..8 members, my_index inside the group");
    rsdsfm_peer_disconnect(ctx);
    if (!ctx->mailbox) RS_TRY(peer_mailbox(ctx));
    for (int g = 0; g < n_members; ++g) {
        if (g == my_index) { ctx->peer_mail[g] = ctx->mailbox; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + (size_t)g * RSDSFM_PEER_HANDLE_BYTES, sizeof h);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { rsdsfm_peer_disconnect(ctx); return fail(ctx, RSDSFM_ERR_CUDA, "cudaIpcOpenMemHandle (peer mailbox)", e); }
        ctx->peer_mail[g] = p; ctx->peer_ipc[g] = true;
    }
    ctx->n_peers = n_members; ctx->my_peer = my_index; ctx->peer_epoch = 0;
    return RSDSFM_OK;
}

int rsdsfm_peer_connect_local(rsdsfm_ctx **members, int n_members)
{
    if (!members || n_members < 1 || n_members > 8) return RSDSFM_ERR_ARG;
    for (int g = 0; g < n_members; ++g) {
        if (!members[g]) return RSDSFM_ERR_ARG;
        rsdsfm_peer_disconnect(members[g]);
        RS_TRY(peer_mailbox(members[g]));
    }
    for (int g = 0; g < n_members; ++g) {
        rsdsfm_ctx *c = members[g];
        RS_CUDA(c, cudaSetDevice(c->device));
        for (int o = 0; o < n_members; ++o) {
            if (o != g && members[o]->device != c->device) {
                int can = 0;
                RS_CUDA(c, cudaDeviceCanAccessPeer(&can, c->device, members[o]->device));
                if (!can) return fail(c, RSDSFM_ERR_CUDA, "rsdsfm_peer_connect_local: no peer access between the devices");
                cudaError_t e = cudaDeviceEnablePeerAccess(members[o]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(c, RSDSFM_ERR_CUDA, "cudaDeviceEnablePeerAccess", e);
                cudaGetLastError();
            }
            c->peer_mail[o] = members[o]->mailbox;
        }
        c->n_peers = n_members; c->my_peer = g; c->peer_epoch = 0;
    }
    return RSDSFM_OK;
}

const char *rsdsfm_last_error(rsdsfm_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int rsdsfm_synchronize(rsdsfm_ctx *ctx)
{
    RS_ENTER(ctx);
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RSDSFM_OK;
}

long long rsdsfm_launch_count(rsdsfm_ctx *ctx) { return ctx ? ctx->launches : 0; }

int rsdsfm_profile_enable(rsdsfm_ctx *ctx, int on)
{
    if (!ctx) return RSDSFM_ERR_ARG;
    ctx->profile = on != 0;
    for (double &p : ctx->prof) p = 0.0;
    for (double &p : ctx->prof_detail) p = 0.0;
    return RSDSFM_OK;
}

int rsdsfm_profile_read(rsdsfm_ctx *ctx, double *out8)
{
    if (!ctx || !out8) return RSDSFM_ERR_ARG;
    for (int j = 0; j < 8; ++j) out8[j] = ctx->prof[j];
    return RSDSFM_OK;
}

int rsdsfm_profile_detail(rsdsfm_ctx *ctx, double *out8)
{
    if (!ctx || !out8) return RSDSFM_ERR_ARG;
    for (int j = 0; j < 8; ++j) out8[j] = ctx->prof_detail[j];
    return RSDSFM_OK;
}

int rsdsfm_flatten(rsdsfm_ctx *ctx, int mem, const double *flow_img, int rows, int cols, const double *K4, double gamma,
                   double flow_threshold, double *coord, double *flow, double *coord_px, double *flow_px,
                   int32_t *pixel_index, int *n_out)
{
    RS_ENTER(ctx);
    if (!flow_img || !coord || !flow || !coord_px || !flow_px || !n_out || rows <= 0 || cols <= 0)
        return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_flatten: bad argument");
    const size_t tot = (size_t)rows * cols;
    const void *d_img = nullptr;
    void *d_c = nullptr, *d_f = nullptr, *d_cp = nullptr, *d_fp = nullptr, *d_pi = nullptr;
    RS_TRY(stage_in(ctx, mem, 0, flow_img, sizeof(double) * 2 * tot, &d_img));
    RS_TRY(stage_out_reserve(ctx, mem, 1, coord, sizeof(double) * 2 * tot, &d_c));
    RS_TRY(stage_out_reserve(ctx, mem, 2, flow, sizeof(double) * 2 * tot, &d_f));
    RS_TRY(stage_out_reserve(ctx, mem, 3, coord_px, sizeof(double) * 2 * tot, &d_cp));
    RS_TRY(stage_out_reserve(ctx, mem, 4, flow_px, sizeof(double) * 2 * tot, &d_fp));
    RS_TRY(stage_out_reserve(ctx, mem, 5, pixel_index, sizeof(int32_t) * tot, &d_pi));
    RS_TRY(flatten_device(ctx, (const double *)d_img, rows, cols, K4, gamma, flow_threshold, (double *)d_c, (double *)d_f,
                          (double *)d_cp, (double *)d_fp, (int32_t *)d_pi, n_out));
    RS_TRY(stage_out(ctx, mem, coord, d_c, sizeof(double) * 2 * tot));
    RS_TRY(stage_out(ctx, mem, flow, d_f, sizeof(double) * 2 * tot));
    RS_TRY(stage_out(ctx, mem, coord_px, d_cp, sizeof(double) * 2 * tot));
    RS_TRY(stage_out(ctx, mem, flow_px, d_fp, sizeof(double) * 2 * tot));
    RS_TRY(stage_out(ctx, mem, pixel_index, d_pi, sizeof(int32_t) * tot));
    return finish_host_call(ctx, mem);
}

int rsdsfm_alpha(rsdsfm_ctx *ctx, int mem, const double *flow_px, const double *q_px, int n, double h, double gamma,
                 double *alpha, double *alpha_k)
{
    RS_ENTER(ctx);
    if (n < 0 || !flow_px || (alpha_k && !q_px)) return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_alpha: bad argument");
    const void *d_f = nullptr, *d_q = nullptr;
    void *d_a = nullptr, *d_ak = nullptr;
    RS_TRY(stage_in(ctx, mem, 0, flow_px, sizeof(double) * 2 * (size_t)n, &d_f));
    RS_TRY(stage_in(ctx, mem, 1, q_px, sizeof(double) * 2 * (size_t)n, &d_q));
    RS_TRY(stage_out_reserve(ctx, mem, 2, alpha, sizeof(double) * (size_t)n, &d_a));
    RS_TRY(stage_out_reserve(ctx, mem, 3, alpha_k, sizeof(double) * (size_t)n, &d_ak));
    RS_TRY(alpha_device(ctx, (const double *)d_f, (const double *)d_q, n, h, gamma, (double *)d_a, (double *)d_ak));
    RS_TRY(stage_out(ctx, mem, alpha, d_a, sizeof(double) * (size_t)n));
    RS_TRY(stage_out(ctx, mem, alpha_k, d_ak, sizeof(double) * (size_t)n));
    return finish_host_call(ctx, mem);
}

int rsdsfm_solve9(const double *q9, const double *u9, const double *alpha9, const double *alpha_k9, int use_alpha_k,
                  double *out7)
{
    if (!q9 || !u9 || !alpha9 || !alpha_k9 || !out7) return RSDSFM_ERR_ARG;
    s9::calculate_velocities(q9, u9, alpha9, alpha_k9, use_alpha_k != 0, out7);
    return RSDSFM_OK;
}

int rsdsfm_ransac_score(rsdsfm_ctx *ctx, int mem, const double *q, const double *u, const double *alpha,
                        const double *alpha_k, int n, const double *hyps, int H, double tolerance, int *counts,
                        double *sumerr, int *best_idx, uint8_t *mask_best, double *inv_depth_best)
{
    RS_ENTER(ctx);
    if (n < 0 || H < 0 || !hyps || !best_idx || (n > 0 && (!q || !u || !alpha || !alpha_k)))
        return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_ransac_score: bad argument");
    const void *d_q = nullptr, *d_u = nullptr, *d_a = nullptr, *d_ak = nullptr;
    void *d_m = nullptr, *d_id = nullptr;
    const size_t nn = (size_t)n;
    RS_TRY(stage_in(ctx, mem, 0, q, sizeof(double) * 2 * nn, &d_q));
    RS_TRY(stage_in(ctx, mem, 1, u, sizeof(double) * 2 * nn, &d_u));
    RS_TRY(stage_in(ctx, mem, 2, alpha, sizeof(double) * nn, &d_a));
    RS_TRY(stage_in(ctx, mem, 3, alpha_k, sizeof(double) * nn, &d_ak));
    RS_TRY(stage_out_reserve(ctx, mem, 4, mask_best, nn, &d_m));
    RS_TRY(stage_out_reserve(ctx, mem, 5, inv_depth_best, sizeof(double) * nn, &d_id));
    RS_TRY(ransac_score_device(ctx, (const double *)d_q, (const double *)d_u, (const double *)d_a, (const double *)d_ak, n,
                               hyps, H, tolerance, counts, sumerr, best_idx, (uint8_t *)d_m, (double *)d_id));
    RS_TRY(stage_out(ctx, mem, mask_best, d_m, nn));
    RS_TRY(stage_out(ctx, mem, inv_depth_best, d_id, sizeof(double) * nn));
    return finish_host_call(ctx, mem);
}

int rsdsfm_ransac(rsdsfm_ctx *ctx, int mem, const double *q, const double *u, const double *alpha, const double *alpha_k,
                  int n, int use_alpha_k, const int32_t *samples, int H, double tolerance, int *counts, double *sumerr,
                  int *best_idx, double *best7, uint8_t *mask_best, double *inv_depth_best, double *hyps_out)
{
    RS_ENTER(ctx);
    if (n <= 0 || H <= 0 || !samples || !best_idx || !q || !u || !alpha || !alpha_k)
        return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_ransac: bad argument");
    const void *d_q = nullptr, *d_u = nullptr, *d_a = nullptr, *d_ak = nullptr;
    void *d_m = nullptr, *d_id = nullptr;
    const size_t nn = (size_t)n;
    RS_TRY(stage_in(ctx, mem, 0, q, sizeof(double) * 2 * nn, &d_q));
    RS_TRY(stage_in(ctx, mem, 1, u, sizeof(double) * 2 * nn, &d_u));
    RS_TRY(stage_in(ctx, mem, 2, alpha, sizeof(double) * nn, &d_a));
    RS_TRY(stage_in(ctx, mem, 3, alpha_k, sizeof(double) * nn, &d_ak));
    RS_TRY(stage_out_reserve(ctx, mem, 4, mask_best, nn, &d_m));
    RS_TRY(stage_out_reserve(ctx, mem, 5, inv_depth_best, sizeof(double) * nn, &d_id));
    std::vector<double> hyps((size_t)H * 7);
    RS_TRY(ransac_fit_device(ctx, (const double *)d_q, (const double *)d_u, (const double *)d_a, (const double *)d_ak, n,
                             use_alpha_k, samples, H, hyps.data()));
    if (hyps_out) memcpy(hyps_out, hyps.data(), sizeof(double) * hyps.size());
    RS_TRY(ransac_score_device(ctx, (const double *)d_q, (const double *)d_u, (const double *)d_a, (const double *)d_ak, n,
                               hyps.data(), H, tolerance, counts, sumerr, best_idx, (uint8_t *)d_m, (double *)d_id));
    if (best7) {
        for (int j = 0; j < 7; ++j) best7[j] = (*best_idx >= 0) ? hyps[(size_t)*best_idx * 7 + j] : 0.0;
    }
    RS_TRY(stage_out(ctx, mem, mask_best, d_m, nn));
    RS_TRY(stage_out(ctx, mem, inv_depth_best, d_id, sizeof(double) * nn));
    return finish_host_call(ctx, mem);
}

int rsdsfm_gather_inliers(rsdsfm_ctx *ctx, int mem, const double *q, const double *alpha, const double *alpha_k, int n,
                          const uint8_t *mask, const double *inv_depth, double *inliers3, double *alpha_in,
                          double *alpha_k_in, int32_t *index_in, int *m_out)
{
    RS_ENTER(ctx);
    if (n < 0 || !m_out || (n > 0 && (!q || !alpha || !alpha_k || !mask || !inv_depth || !inliers3 || !alpha_in || !alpha_k_in)))
        return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_gather_inliers: bad argument");
    const size_t nn = (size_t)n;
    const void *d_q = nullptr, *d_a = nullptr, *d_ak = nullptr, *d_m = nullptr, *d_id = nullptr;
    void *d_i3 = nullptr, *d_ai = nullptr, *d_aki = nullptr, *d_ix = nullptr;
    RS_TRY(stage_in(ctx, mem, 0, q, sizeof(double) * 2 * nn, &d_q));
    RS_TRY(stage_in(ctx, mem, 1, alpha, sizeof(double) * nn, &d_a));
    RS_TRY(stage_in(ctx, mem, 2, alpha_k, sizeof(double) * nn, &d_ak));
    RS_TRY(stage_in(ctx, mem, 3, mask, nn, &d_m));
    RS_TRY(stage_in(ctx, mem, 4, inv_depth, sizeof(double) * nn, &d_id));
    RS_TRY(stage_out_reserve(ctx, mem, 5, inliers3, sizeof(double) * 3 * nn, &d_i3));
    RS_TRY(stage_out_reserve(ctx, mem, 6, alpha_in, sizeof(double) * nn, &d_ai));
    RS_TRY(stage_out_reserve(ctx, mem, 7, alpha_k_in, sizeof(double) * nn, &d_aki));
    RS_TRY(stage_out_reserve(ctx, mem, 8, index_in, sizeof(int32_t) * nn, &d_ix));
    RS_TRY(gather_inliers_device(ctx, (const double *)d_q, (const double *)d_a, (const double *)d_ak, n, (const uint8_t *)d_m,
                                 (const double *)d_id, (double *)d_i3, (double *)d_ai, (double *)d_aki, (int32_t *)d_ix, m_out));
    const size_t m = (size_t)*m_out;
    RS_TRY(stage_out(ctx, mem, inliers3, d_i3, sizeof(double) * 3 * m));
    RS_TRY(stage_out(ctx, mem, alpha_in, d_ai, sizeof(double) * m));
    RS_TRY(stage_out(ctx, mem, alpha_k_in, d_aki, sizeof(double) * m));
    RS_TRY(stage_out(ctx, mem, index_in, d_ix, sizeof(int32_t) * m));
    return finish_host_call(ctx, mem);
}

int rsdsfm_estimate_inverse_depths(rsdsfm_ctx *ctx, int mem, const double *coord, const double *flow, int n,
                                   const double *v, const double *w, double k, const double *alpha,
                                   const double *alpha_k, double *inv_depth, rsdsfm_lm_summary *summary)
{
    RS_ENTER(ctx);
    if (n < 0 || !v || !w || (n > 0 && (!coord || !flow || !alpha || !alpha_k || !inv_depth)))
        return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_estimate_inverse_depths: bad argument");
    const size_t nn = (size_t)n;
    const void *d_c = nullptr, *d_f = nullptr, *d_a = nullptr, *d_ak = nullptr;
    void *d_o = nullptr;
    RS_TRY(stage_in(ctx, mem, 0, coord, sizeof(double) * 2 * nn, &d_c));
    RS_TRY(stage_in(ctx, mem, 1, flow, sizeof(double) * 2 * nn, &d_f));
    RS_TRY(stage_in(ctx, mem, 2, alpha, sizeof(double) * nn, &d_a));
    RS_TRY(stage_in(ctx, mem, 3, alpha_k, sizeof(double) * nn, &d_ak));
    RS_TRY(stage_out_reserve(ctx, mem, 4, inv_depth, sizeof(double) * nn, &d_o));
    RS_TRY(estimate_inverse_depths_device(ctx, (const double *)d_c, (const double *)d_f, n, v, w, k, (const double *)d_a,
                                          (const double *)d_ak, (double *)d_o, summary));
    RS_TRY(stage_out(ctx, mem, inv_depth, d_o, sizeof(double) * nn));
    return finish_host_call(ctx, mem);
}

int rsdsfm_refine(rsdsfm_ctx *ctx, int mem, const double *flow, const double *inliers3, const double *alpha,
                  const double *alpha_k, int m, double *v, double *w, double *k, int const_acceleration,
                  const int32_t *flow_index, const rsdsfm_lm_options *opts, double *z_out, rsdsfm_lm_summary *summary)
{
    RS_ENTER(ctx);
    if (m < 0 || !v || !w || !k || (m > 0 && (!flow || !inliers3 || !alpha || !alpha_k || !z_out)))
        return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_refine: bad argument");
    const size_t mm = (size_t)m;
    const void *d_f = nullptr, *d_i = nullptr, *d_a = nullptr, *d_ak = nullptr, *d_fi = nullptr;
    void *d_z = nullptr;
    // the reference reads flow(:, i) for i < m of whatever array it is handed; with a repaired
    // pairing the caller's array must cover max(flow_index)+1 columns -- host callers pass that
    // length through the device path only (flow_index needs RSDSFM_DEVICE or m columns).
    RS_TRY(stage_in(ctx, mem, 0, flow, sizeof(double) * 2 * mm, &d_f));
    RS_TRY(stage_in(ctx, mem, 1, inliers3, sizeof(double) * 3 * mm, &d_i));
    RS_TRY(stage_in(ctx, mem, 2, alpha, sizeof(double) * mm, &d_a));
    RS_TRY(stage_in(ctx, mem, 3, alpha_k, sizeof(double) * mm, &d_ak));
    RS_TRY(stage_in(ctx, mem, 4, flow_index, sizeof(int32_t) * mm, &d_fi));
    RS_TRY(stage_out_reserve(ctx, mem, 5, z_out, sizeof(double) * mm, &d_z));
    if (mem == RSDSFM_HOST && flow_index)
        return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_refine: flow_index requires RSDSFM_DEVICE buffers (flow length is not known)");
    RS_TRY(refine_device(ctx, (const double *)d_f, (const double *)d_i, (const double *)d_a, (const double *)d_ak, m, v, w, k,
                         const_acceleration, (const int32_t *)d_fi, opts, (double *)d_z, summary));
    RS_TRY(stage_out(ctx, mem, z_out, d_z, sizeof(double) * mm));
    return finish_host_call(ctx, mem);
}

int rsdsfm_depth_glue(rsdsfm_ctx *ctx, int mem, double *inliers3, int m, double *v, const double *K4, int rows, int cols,
                      double z_min_init, int layout, double *depth_map, uint8_t *depth_img)
{
    RS_ENTER(ctx);
    if (m < 0 || rows <= 0 || cols <= 0 || !K4 || !depth_map || !v || (m > 0 && !inliers3))
        return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_depth_glue: bad argument");
    const size_t mm = (size_t)m, tot = (size_t)rows * cols;
    const void *d_in = nullptr;
    void *d_dm = nullptr, *d_di = nullptr;
    RS_TRY(stage_in(ctx, mem, 0, inliers3, sizeof(double) * 3 * mm, &d_in));
    RS_TRY(stage_out_reserve(ctx, mem, 1, depth_map, sizeof(double) * tot, &d_dm));
    RS_TRY(stage_out_reserve(ctx, mem, 2, depth_img, tot, &d_di));
    RS_TRY(ensure(ctx, ctx->misc, 256));
    double *stats = (double *)ctx->misc.p;
    double *inl = (double *)d_in;
    RS_TRY(glue_device(ctx, inl + 2, 3, inl, 3, m, K4, rows, cols, z_min_init, layout, (double *)d_dm, (uint8_t *)d_di, stats));
    if (m > 0) {
        RS_TRY(ensure_pinned(ctx, 1024));
        RS_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, stats, sizeof(double) * 8, cudaMemcpyDeviceToHost, ctx->stream));
        RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (((double *)ctx->pinned)[3] < 0) { v[0] *= -1.0; v[1] *= -1.0; v[2] *= -1.0; }
    }
    RS_TRY(stage_out(ctx, mem, inliers3, d_in, sizeof(double) * 3 * mm));
    RS_TRY(stage_out(ctx, mem, depth_map, d_dm, sizeof(double) * tot));
    RS_TRY(stage_out(ctx, mem, depth_img, d_di, tot));
    return finish_host_call(ctx, mem);
}

int rsdsfm_set_relative_pose(const double *v, const double *w, double k, double gamma, int rows, double *R, double *t)
{
    if (!v || !w || !R || !t || rows <= 0) return RSDSFM_ERR_ARG;
    // RsFrame::setRelativePose, rsframe.cc:771-800 (host restatement; the device pipeline uses
    // k_set_relative_pose in rectify.cu, same operation order)
    const double skew[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    for (int i = 0; i < rows; ++i) {
        double *Ri = R + 9 * (size_t)i, *ti = t + 3 * (size_t)i;
        const double beta_1 = (i == 0) ? 0.0
            : (gamma * i / rows + 0.5 * k * (gamma * gamma * i * i) / (rows * rows)) * (2.0 / (2.0 + k));
        for (int a = 0; a < 9; ++a) Ri[a] = ((a % 4 == 0) ? 1.0 : 0.0) + ((i == 0) ? 0.0 : beta_1 * skew[a]);
        for (int a = 0; a < 3; ++a) ti[a] = (i == 0) ? 0.0 : 0.0 + beta_1 * v[a];
    }
    return RSDSFM_OK;
}

int rsdsfm_backproject(rsdsfm_ctx *ctx, int mem, const uint8_t *image, const double *depth, int layout, int rows, int cols,
                       const double *K4, const double *R, const double *t, int gs_mode, uint8_t *gs_out, float *coords3d)
{
    RS_ENTER(ctx);
    if (rows <= 0 || cols <= 0 || !image || !depth || !K4 || !R || !t || !gs_out)
        return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_backproject: bad argument");
    const size_t tot = (size_t)rows * cols;
    const void *d_img = nullptr, *d_dep = nullptr;
    void *d_gs = nullptr, *d_c3 = nullptr;
    RS_TRY(stage_in(ctx, mem, 0, image, tot * 3, &d_img));
    RS_TRY(stage_in(ctx, mem, 1, depth, sizeof(double) * tot, &d_dep));
    RS_TRY(stage_out_reserve(ctx, mem, 2, gs_out, tot * 3, &d_gs));
    RS_TRY(stage_out_reserve(ctx, mem, 3, coords3d, sizeof(float) * 3 * tot, &d_c3));
    RS_TRY(ensure(ctx, ctx->poses, sizeof(double) * 12 * (size_t)rows));
    double *dR = (double *)ctx->poses.p, *dt = dR + 9 * (size_t)rows;
    RS_CUDA(ctx, cudaMemcpyAsync(dR, R, sizeof(double) * 9 * (size_t)rows, cudaMemcpyHostToDevice, ctx->stream));
    RS_CUDA(ctx, cudaMemcpyAsync(dt, t, sizeof(double) * 3 * (size_t)rows, cudaMemcpyHostToDevice, ctx->stream));
    RS_TRY(backproject_device(ctx, (const uint8_t *)d_img, (const double *)d_dep, layout, rows, cols, K4, dR, dt, gs_mode,
                              (uint8_t *)d_gs, (float *)d_c3));
    RS_TRY(stage_out(ctx, mem, gs_out, d_gs, tot * 3));
    RS_TRY(stage_out(ctx, mem, coords3d, d_c3, sizeof(float) * 3 * tot));
    // R, t are host arrays: the copies above must have been consumed before the caller reuses them
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RSDSFM_OK;
}

int rsdsfm_fill_cracks(rsdsfm_ctx *ctx, int mem, const uint8_t *in, int rows, int cols, unsigned offset, uint8_t *out)
{
    RS_ENTER(ctx);
    if (rows <= 0 || cols <= 0 || !in || !out) return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_fill_cracks: bad argument");
    const size_t tot = (size_t)rows * cols;
    const void *d_in = nullptr;
    void *d_out = nullptr;
    RS_TRY(stage_in(ctx, mem, 0, in, tot * 3, &d_in));
    RS_TRY(stage_out_reserve(ctx, mem, 1, out, tot * 3, &d_out));
    RS_TRY(fill_cracks_device(ctx, (const uint8_t *)d_in, rows, cols, offset, (uint8_t *)d_out));
    RS_TRY(stage_out(ctx, mem, out, d_out, tot * 3));
    return finish_host_call(ctx, mem);
}

int rsdsfm_relocate_pose(const double *R, const double *t, int rows, double *R_out, double *t_out)
{
    if (!R || !t || !R_out || !t_out || rows <= 0) return RSDSFM_ERR_ARG;
    // RsFrame::relocatePose, rsframe.cc:953-967: scanline 0 keeps its pose; for i >= 1
    // t_i -= t_0 and R_i = R_0^-1 R_i with Eigen's 3x3 inverse (cofactors / determinant)
    const double *m = R;
    auto cof = [&](int i, int j) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
    };
    const double det = cof(0, 0) * m[0] + cof(1, 0) * m[3] + cof(2, 0) * m[6];
    const double invdet = 1.0 / det;
    double inv[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) inv[r * 3 + c] = cof(c, r) * invdet;
    const double t0[3] = {t[0], t[1], t[2]};
    for (int a = 0; a < 9; ++a) R_out[a] = R[a];
    for (int a = 0; a < 3; ++a) t_out[a] = t[a];
    for (int i = 1; i < rows; ++i) {
        const double *Ri = R + 9 * (size_t)i;
        double Rn[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Rn[r * 3 + c] = inv[r * 3 + 0] * Ri[0 * 3 + c] + inv[r * 3 + 1] * Ri[1 * 3 + c] + inv[r * 3 + 2] * Ri[2 * 3 + c];
        for (int a = 0; a < 9; ++a) R_out[9 * (size_t)i + a] = Rn[a];
        for (int a = 0; a < 3; ++a) t_out[3 * (size_t)i + a] = t[3 * (size_t)i + a] - t0[a];
    }
    return RSDSFM_OK;
}

int rsdsfm_reprojection_error(rsdsfm_ctx *ctx, int mem, const float *coords3d, const double *unproj_x, const double *unproj_y,
                              const double *unproj_z, const double *R_gt, const double *t_gt, const double *depth_est,
                              int layout, int rows, int cols, const double *K4, double max_norm, double *mean_error,
                              double *mean_scale, int *num_outliers, int *points_used, uint8_t *error_image,
                              double *gt_depth_map)
{
    RS_ENTER(ctx);
    if (rows <= 0 || cols <= 0 || !coords3d || !unproj_x || !unproj_y || !unproj_z || !R_gt || !t_gt || !depth_est || !K4 ||
        !mean_error)
        return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_reprojection_error: bad argument");
    const size_t tot = (size_t)rows * cols;
    const void *d_c = nullptr, *d_x = nullptr, *d_y = nullptr, *d_z = nullptr, *d_de = nullptr;
    void *d_img = nullptr, *d_gd = nullptr;
    RS_TRY(stage_in(ctx, mem, 0, coords3d, sizeof(float) * 3 * tot, &d_c));
    RS_TRY(stage_in(ctx, mem, 1, unproj_x, sizeof(double) * tot, &d_x));
    RS_TRY(stage_in(ctx, mem, 2, unproj_y, sizeof(double) * tot, &d_y));
    RS_TRY(stage_in(ctx, mem, 3, unproj_z, sizeof(double) * tot, &d_z));
    RS_TRY(stage_in(ctx, mem, 4, depth_est, sizeof(double) * tot, &d_de));
    RS_TRY(stage_out_reserve(ctx, mem, 5, error_image, tot, &d_img));
    RS_TRY(stage_out_reserve(ctx, mem, 6, gt_depth_map, sizeof(double) * tot, &d_gd));
    // scanline poses: original + relocated, interleaved per row, staged through pinned memory
    RS_TRY(ensure_pinned(ctx, sizeof(double) * 24 * (size_t)rows + 256));
    RS_TRY(ensure(ctx, ctx->pipe[15], sizeof(double) * 24 * (size_t)rows));
    RS_TRY(ensure(ctx, ctx->pipe[14], sizeof(float) * 3 * tot));
    RS_TRY(ensure(ctx, ctx->misc, 256));
    {
        std::vector<double> Rr((size_t)rows * 9), tr((size_t)rows * 3);
        rsdsfm_relocate_pose(R_gt, t_gt, rows, Rr.data(), tr.data());
        double *hp = (double *)ctx->pinned;
        for (int i = 0; i < rows; ++i) {
            double *o = hp + 24 * (size_t)i;
            for (int a = 0; a < 9; ++a) { o[a] = R_gt[9 * (size_t)i + a]; o[12 + a] = Rr[9 * (size_t)i + a]; }
            for (int a = 0; a < 3; ++a) { o[9 + a] = t_gt[3 * (size_t)i + a]; o[21 + a] = tr[3 * (size_t)i + a]; }
        }
        RS_CUDA(ctx, cudaMemcpyAsync(ctx->pipe[15].p, hp, sizeof(double) * 24 * (size_t)rows, cudaMemcpyHostToDevice, ctx->stream));
    }
    double *sums = (double *)ctx->misc.p;
    RS_TRY(reproj_device(ctx, (const float *)d_c, (const double *)d_x, (const double *)d_y, (const double *)d_z,
                         (const double *)ctx->pipe[15].p, (const double *)d_de, layout, rows, cols, K4, max_norm,
                         (float *)ctx->pipe[14].p, (uint8_t *)d_img, (double *)d_gd, sums));
    RS_TRY(stage_out(ctx, mem, error_image, d_img, tot));
    RS_TRY(stage_out(ctx, mem, gt_depth_map, d_gd, sizeof(double) * tot));
    double *hs = (double *)((char *)ctx->pinned + sizeof(double) * 24 * (size_t)rows);
    RS_CUDA(ctx, cudaMemcpyAsync(hs, sums, sizeof(double) * 5, cudaMemcpyDeviceToHost, ctx->stream));
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *mean_error = hs[3] * 1.0 / hs[4];                                  // camera.cc:690
    if (mean_scale) *mean_scale = hs[0] / hs[1];
    if (num_outliers) *num_outliers = (int)hs[2];
    if (points_used) *points_used = (int)hs[4];
    return RSDSFM_OK;
}

int rsdsfm_true_flow(rsdsfm_ctx *ctx, int mem, const double *unproj_x, const double *unproj_y, const double *unproj_z,
                     const double *R2, const double *t2, int layout, int rows, int cols, const double *K4, double *flow)
{
    RS_ENTER(ctx);
    if (rows <= 0 || cols <= 0 || !unproj_x || !unproj_y || !unproj_z || !R2 || !t2 || !K4 || !flow)
        return fail(ctx, RSDSFM_ERR_ARG, "rsdsfm_true_flow: bad argument");
    const size_t tot = (size_t)rows * cols;
    const void *d_x = nullptr, *d_y = nullptr, *d_z = nullptr;
    void *d_f = nullptr;
    RS_TRY(stage_in(ctx, mem, 0, unproj_x, sizeof(double) * tot, &d_x));
    RS_TRY(stage_in(ctx, mem, 1, unproj_y, sizeof(double) * tot, &d_y));
    RS_TRY(stage_in(ctx, mem, 2, unproj_z, sizeof(double) * tot, &d_z));
    RS_TRY(stage_out_reserve(ctx, mem, 3, flow, sizeof(double) * 2 * tot, &d_f));
    RS_TRY(ensure_pinned(ctx, sizeof(double) * 12 * (size_t)rows));
    RS_TRY(ensure(ctx, ctx->pipe[15], sizeof(double) * 24 * (size_t)rows));
    double *hp = (double *)ctx->pinned;
    for (int i = 0; i < rows; ++i) {
        for (int a = 0; a < 9; ++a) hp[12 * (size_t)i + a] = R2[9 * (size_t)i + a];
        for (int a = 0; a < 3; ++a) hp[12 * (size_t)i + 9 + a] = t2[3 * (size_t)i + a];
    }
    RS_CUDA(ctx, cudaMemcpyAsync(ctx->pipe[15].p, hp, sizeof(double) * 12 * (size_t)rows, cudaMemcpyHostToDevice, ctx->stream));
    RS_TRY(true_flow_device(ctx, (const double *)d_x, (const double *)d_y, (const double *)d_z, (const double *)ctx->pipe[15].p,
                            layout, rows, cols, K4, (double *)d_f));
    RS_TRY(stage_out(ctx, mem, flow, d_f, sizeof(double) * 2 * tot));
    // R2, t2 were staged through the context's pinned buffer: the copy must have been consumed before it is reused
    RS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RSDSFM_OK;
}

int rsdsfm_host_alloc(size_t bytes, int write_combined, void **out)
{
    if (!out || bytes == 0) return RSDSFM_ERR_ARG;
    *out = nullptr;
    const cudaError_t e = cudaHostAlloc(out, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
    if (e != cudaSuccess) return fail(nullptr, RSDSFM_ERR_NOMEM, "cudaHostAlloc", e);
    return RSDSFM_OK;
}

int rsdsfm_host_free(void *p)
{
    if (!p) return RSDSFM_OK;
    return cudaFreeHost(p) == cudaSuccess ? RSDSFM_OK : RSDSFM_ERR_CUDA;
}

}  // extern "C"
