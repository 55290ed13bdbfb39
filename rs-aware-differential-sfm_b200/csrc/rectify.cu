// rectify.cu -- sign fix + depth raster glue, per-scanline relative poses, the global-shutter
// rectification splat and the crack fill.  Compiled with --fmad=false: the per-pixel projection
// mirrors the reference's operation order so the integer splat targets are reproducible.
//
//   a10  main.cc:466-509 / errorMeasure.cpp:162-210      k_glue_* kernels
//   a12  RsFrame::setRelativePose  rsframe.cc:771-800      k_set_relative_pose
//   a13  planeToSpace / cameraToWorldFrame / worldToCameraFrame / spaceToPlane  rsframe.cc:629-736
//   a14  RsFrame::backProject / backProjectGs  rsframe.cc:803-878   k_splat_vote + k_splat_gather
//   a15  Camera::interpolateCrackyImage  camera.cc:694-774  k_fill_cracks
//
// The reference splats sequentially in raster order, so on a collision the source pixel with the
// highest raster index wins (Q15).  Here every source pixel votes with atomicMax(raster index+1)
// into a per-target winner map, and a second kernel gathers the winning colours: deterministic
// and identical to last-writer-wins.
#include "common.cuh"

namespace rsdsfm {

// out[j] = reduce over CTA rows b of partials[b*(ns+nm)+j] for up to 8 columns: warp c handles
// column c; lane l combines rows l, l+32, ... in ascending order, then a butterfly over the lanes
// (fixed order => bit-reproducible for a given number of rows).
__global__ void __launch_bounds__(256) k_final_reduce(const double *__restrict__ partials, int nblocks, int ns, int nm, double *out)
{
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31, nv = ns + nm;
    if (c >= nv) return;
    const bool is_sum = c < ns;
    double v = is_sum ? 0.0 : -INFINITY;
    for (int b = lane; b < nblocks; b += 32) {
        const double x = partials[(size_t)b * nv + c];
        v = is_sum ? v + x : fmax(v, x);
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double y = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_sum ? v + y : fmax(v, y);
    }
    if (lane == 0) out[c] = v;
}

void launch_final_reduce(rsdsfm_ctx *ctx, const double *partials, int nblocks, int ns, int nm, double *out)
{
    k_final_reduce<<<1, 256, 0, ctx->stream>>>(partials, nblocks, ns, nm, out);   // ns + nm <= 8
    ctx->launches++;
}

// ---------------------------------------------------------------- a10: sign fix + depth raster
// sums[0] = sum z, maxima: [0] = max z, [1] = max(-z)  (=> min z)
__global__ void __launch_bounds__(kThreads) k_glue_reduce(const double *__restrict__ z, int zs, int m, double *partials)
{
    double s[1] = {0.0};
    double mx[2] = {-INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const double zi = z[(size_t)i * zs];
        s[0] += zi;
        mx[0] = fmax(mx[0], zi);
        mx[1] = fmax(mx[1], -zi);
    }
    block_reduce_store<1, 2>(s, mx, partials);
}

// rows: nrows x {sum z, max z, max -z} (one row per CTA of the producing kernel).  Reduces them in a
// fixed order (warp c = column c; lane l takes rows l, l+32, ...; shuffle tree) and derives
// stats (device, 8 doubles): [0] sum z, [1] max z, [2] max -z, [3] sign (+1/-1), [4] z_min, [5] z_max
// after the sign fix (z_min starts at z_min_init, z_max at 0 like the reference, main.cc:481-489).
__global__ void __launch_bounds__(96) k_glue_stats(const double *__restrict__ rows, int nrows, double *stats, int m,
                                                   double z_min_init)
{
    __shared__ double red[3];
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double v = (c == 0) ? 0.0 : -INFINITY;
    for (int b = lane; b < nrows; b += 32) {
        const double x = rows[(size_t)b * 3 + c];
        v = (c == 0) ? v + x : fmax(v, x);
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double y = __shfl_xor_sync(0xffffffffu, v, o);
        v = (c == 0) ? v + y : fmax(v, y);
    }
    if (lane == 0) red[c] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        const double z_mean = red[0] * 1.0 / m;
        const double sign = (z_mean < 0) ? -1.0 : 1.0;
        const double zmax = sign > 0 ? red[1] : red[2];
        const double zmin = sign > 0 ? -red[2] : -red[1];
        stats[0] = red[0]; stats[1] = red[1]; stats[2] = red[2];
        stats[3] = sign;
        stats[4] = fmin(z_min_init, zmin);
        stats[5] = fmax(0.0, zmax);
    }
}

__device__ __forceinline__ bool to_int_trunc(double a, int &out)
{
    if (!(fabs(a) < 2147483648.0)) return false;   // NaN / Inf / overflow: rejected (x86 gives INT_MIN)
    out = (int)a;
    return true;
}

// z (stride zs doubles, e.g. 3 for the z row of an Array3Xd) is sign-fixed in place; xy gives the
// normalised coordinates (stride 3 when interleaved in inliers3).
__global__ void k_glue_raster(double *z, int zs, const double *__restrict__ xyz, int xs, int m,
                              const double *__restrict__ stats, double fx, double fy, double cx, double cy, int rows,
                              int cols, int layout, double *depth_map, uint8_t *depth_img)
{
    const double sign = stats[3], z_min = stats[4], z_max = stats[5];
    const double multiplier = 244.0 / (z_max - z_min);
    // four inliers per thread and round: all loads first (independent), then the scatters
    constexpr int kU = 4;
    const int stride = gridDim.x * blockDim.x;
    for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < m; i0 += kU * stride) {
        double zz4[kU], xx4[kU], yy4[kU];
#pragma unroll
        for (int r = 0; r < kU; ++r) {
            const int i = i0 + r * stride;
            const bool in = i < m;
            zz4[r] = in ? z[(size_t)i * zs] : 0.0;
            xx4[r] = in ? xyz[(size_t)i * xs] : 0.0;
            yy4[r] = in ? xyz[(size_t)i * xs + 1] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < kU; ++r) {
            const int i = i0 + r * stride;
            if (i >= m) break;
            double zi = zz4[r];
            if (sign < 0) { zi = zi * -1.0; z[(size_t)i * zs] = zi; }
            const double xd = fx * xx4[r] + cx + 0.5;
            const double yd = fy * yy4[r] + cy + 0.5;
            int x, y;
            if (!to_int_trunc(xd, x) || !to_int_trunc(yd, y)) continue;
            if (x < 0 || x >= cols || y < 0 || y >= rows) continue;   // UB in the reference (Q3): skipped
            if (depth_img) {
                const double zz = (zi - z_min) * multiplier;
                int zq = 10;
                if (fabs(zz) < 2147483000.0) zq = 10 + (int)zz;
                depth_img[(size_t)y * cols + x] = (uint8_t)zq;
            }
            const size_t idx = (layout == RSDSFM_DEPTH_COLMAJOR) ? ((size_t)y + (size_t)x * rows) : ((size_t)y * cols + x);
            depth_map[idx] = zi;
        }
    }
}

// ---------------------------------------------------------------- a12: per-scanline poses
// motion7 (device): v[3], w[3], k.  stats may be NULL (no sign fix), else stats[3] multiplies v.
__global__ void k_set_relative_pose(const double *__restrict__ motion7, const double *__restrict__ stats, double gamma,
                                    int rows, double *R, double *t)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const double sign = stats ? stats[3] : 1.0;
    double v[3] = {motion7[0], motion7[1], motion7[2]};
    if (sign < 0) { v[0] *= -1.0; v[1] *= -1.0; v[2] *= -1.0; }
    const double w0 = motion7[3], w1 = motion7[4], w2 = motion7[5], k = motion7[6];
    double *Ri = R + 9 * (size_t)i, *ti = t + 3 * (size_t)i;
    if (i == 0) {
        Ri[0] = 1; Ri[1] = 0; Ri[2] = 0; Ri[3] = 0; Ri[4] = 1; Ri[5] = 0; Ri[6] = 0; Ri[7] = 0; Ri[8] = 1;
        ti[0] = 0; ti[1] = 0; ti[2] = 0;
        return;
    }
    const double skew[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};
    const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const double beta_1 = (gamma * i / rows + 0.5 * k * (gamma * gamma * i * i) / (rows * rows)) * (2.0 / (2.0 + k));
    double Rn[9];
    for (int a = 0; a < 9; ++a) Rn[a] = I3[a] + beta_1 * skew[a];
    for (int r = 0; r < 3; ++r)                                   // R0 * R_new with R0 = I (:797)
        for (int c = 0; c < 3; ++c)
            Ri[r * 3 + c] = I3[r * 3 + 0] * Rn[0 * 3 + c] + I3[r * 3 + 1] * Rn[1 * 3 + c] + I3[r * 3 + 2] * Rn[2 * 3 + c];
    for (int a = 0; a < 3; ++a) ti[a] = 0.0 + beta_1 * v[a];
}

// ---------------------------------------------------------------- a13 + a14: splat
struct SplatParams {
    double fx, fy, cx, cy;
    int rows, cols, layout, gs_mode;
};

__global__ void __launch_bounds__(kThreads) k_splat_vote(const uint8_t *__restrict__ image, const double *__restrict__ depth,
                                                         const double *__restrict__ R, const double *__restrict__ t,
                                                         SplatParams P, unsigned int *winner, float *coords3d)
{
    const long long total = (long long)P.rows * P.cols;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(p / P.cols), x = (int)(p - (long long)y * P.cols);
        const uint8_t b = image[3 * p], g = image[3 * p + 1], r = image[3 * p + 2];
        if (coords3d) { coords3d[3 * p] = 0.f; coords3d[3 * p + 1] = 0.f; coords3d[3 * p + 2] = 0.f; }
        if (b == 1 && g == 1 && r == 1) continue;                                   // rsframe.cc:815
        const int s = P.gs_mode ? 0 : y;
        const double *Rs = R + 9 * (size_t)s, *ts = t + 3 * (size_t)s;
        // inverse pose [R^T | -R^T t]  (rsframe.cc:719-733)
        double Rt[9], ti[3];
        for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) Rt[a * 3 + c] = Rs[c * 3 + a];
        for (int a = 0; a < 3; ++a) ti[a] = (-Rt[a * 3 + 0]) * ts[0] + (-Rt[a * 3 + 1]) * ts[1] + (-Rt[a * 3 + 2]) * ts[2];
        // planeToSpace (rsframe.cc:646-665)
        const double nx = ((double)x - P.cx) * 1.0 / P.fx;
        const double ny = ((double)y - P.cy) * 1.0 / P.fy;
        const double z = depth[(P.layout == RSDSFM_DEPTH_COLMAJOR) ? ((size_t)y + (size_t)x * P.rows) : (size_t)p];
        const double Pc[3] = {z * nx, z * ny, z * 1.0};
        double Pw[3], Pg[3];
        for (int a = 0; a < 3; ++a)
            Pw[a] = Rt[a * 3 + 0] * Pc[0] + Rt[a * 3 + 1] * Pc[1] + Rt[a * 3 + 2] * Pc[2] + ti[a] * 1.0;
        for (int a = 0; a < 3; ++a)                                                  // scanline 0 pose (:821)
            Pg[a] = R[a * 3 + 0] * Pw[0] + R[a * 3 + 1] * Pw[1] + R[a * 3 + 2] * Pw[2] + t[a] * 1.0;
        const double u = Pg[0] / Pg[2] * P.fx + P.cx;                                // spaceToPlane (:629-642)
        const double v = Pg[1] / Pg[2] * P.fx + P.cy;                                // y uses f_x too (Q12)
        if (coords3d) { coords3d[3 * p] = (float)Pw[0]; coords3d[3 * p + 1] = (float)Pw[1]; coords3d[3 * p + 2] = (float)Pw[2]; }
        int tx, ty;
        if (!to_int_trunc(u + 0.5, tx) || !to_int_trunc(v + 0.5, ty)) continue;
        if (tx >= 0 && tx < P.cols && ty >= 0 && ty < P.rows)
            atomicMax(&winner[(size_t)ty * P.cols + tx], (unsigned int)(p + 1));
    }
}

// ---- the same vote with everything that does not depend on the pixel taken out of the per-pixel path:
//   k_splat_tables  per scanline the inverse pose [R^T | -R^T t] (12 doubles) and ny = (y - cy) * 1.0 / fy,
//                   per column nx = (x - cx) * 1.0 / fx -- the very operations k_splat_vote performs per pixel, once;
//   k_splat_vote4   one row per CTA row, four consecutive pixels per thread: the frame arrives as three 32-bit
//                   words, the depths and nx as two 16-byte loads each; two IEEE divisions per pixel are left.
// Table layout: inv[rows][12] | ny[rows] | (pad to an even count) | nx[cols].
__global__ void k_splat_tables(const double *__restrict__ R, const double *__restrict__ t, SplatParams P, double *tab)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double *inv = tab, *ny = tab + 12 * (size_t)P.rows, *nx = tab + ((13 * (size_t)P.rows + 1) & ~(size_t)1);
    if (i < P.rows) {
        const double *Rs = R + 9 * (size_t)i, *ts = t + 3 * (size_t)i;
        double Rt[9];
        for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) Rt[a * 3 + c] = Rs[c * 3 + a];
        for (int a = 0; a < 9; ++a) inv[12 * (size_t)i + a] = Rt[a];
        for (int a = 0; a < 3; ++a)
            inv[12 * (size_t)i + 9 + a] = (-Rt[a * 3 + 0]) * ts[0] + (-Rt[a * 3 + 1]) * ts[1] + (-Rt[a * 3 + 2]) * ts[2];
        ny[i] = ((double)i - P.cy) * 1.0 / P.fy;
    }
    if (i < P.cols) nx[i] = ((double)i - P.cx) * 1.0 / P.fx;
}

__global__ void __launch_bounds__(kThreads) k_splat_vote4(const uint8_t *__restrict__ image, const double *__restrict__ depth,
                                                          const double *__restrict__ R, const double *__restrict__ t,
                                                          const double *__restrict__ tab, SplatParams P, unsigned int *winner,
                                                          float *coords3d)
{
    const int y = blockIdx.y;
    const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (x0 >= P.cols) return;                                        // (cols is a multiple of 4: whole quads only)
    const double *inv = tab + 12 * (size_t)(P.gs_mode ? 0 : y);
    const double ny = tab[12 * (size_t)P.rows + y];
    const double *nxt = tab + ((13 * (size_t)P.rows + 1) & ~(size_t)1) + x0;
    double Rt[9], ti[3], R0[9], t0[3];
#pragma unroll
    for (int a = 0; a < 9; ++a) { Rt[a] = __ldg(inv + a); R0[a] = __ldg(R + a); }
#pragma unroll
    for (int a = 0; a < 3; ++a) { ti[a] = __ldg(inv + 9 + a); t0[a] = __ldg(t + a); }
    const size_t p0 = (size_t)y * P.cols + x0;
    const uint3 w = *reinterpret_cast<const uint3 *>(image + 3 * p0);                 // 12 bytes = 4 BGR pixels
    const unsigned int bgr[4] = {w.x & 0xffffffu, (w.x >> 24) | ((w.y & 0xffffu) << 8), (w.y >> 16) | ((w.z & 0xffu) << 16), w.z >> 8};
    const double2 za = *reinterpret_cast<const double2 *>(depth + p0), zb = *reinterpret_cast<const double2 *>(depth + p0 + 2);
    const double2 na = *reinterpret_cast<const double2 *>(nxt), nb = *reinterpret_cast<const double2 *>(nxt + 2);
    const double zs[4] = {za.x, za.y, zb.x, zb.y}, nxs[4] = {na.x, na.y, nb.x, nb.y};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const size_t p = p0 + j;
        if (coords3d) { coords3d[3 * p] = 0.f; coords3d[3 * p + 1] = 0.f; coords3d[3 * p + 2] = 0.f; }
        if (bgr[j] == 0x010101u) continue;                                           // rsframe.cc:815
        const double z = zs[j];
        const double Pc[3] = {z * nxs[j], z * ny, z * 1.0};
        double Pw[3], Pg[3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
            Pw[a] = Rt[a * 3 + 0] * Pc[0] + Rt[a * 3 + 1] * Pc[1] + Rt[a * 3 + 2] * Pc[2] + ti[a] * 1.0;
#pragma unroll
        for (int a = 0; a < 3; ++a)                                                  // scanline 0 pose (:821)
            Pg[a] = R0[a * 3 + 0] * Pw[0] + R0[a * 3 + 1] * Pw[1] + R0[a * 3 + 2] * Pw[2] + t0[a] * 1.0;
        const double u = Pg[0] / Pg[2] * P.fx + P.cx;                                // spaceToPlane (:629-642)
        const double v = Pg[1] / Pg[2] * P.fx + P.cy;                                // y uses f_x too (Q12)
        if (coords3d) { coords3d[3 * p] = (float)Pw[0]; coords3d[3 * p + 1] = (float)Pw[1]; coords3d[3 * p + 2] = (float)Pw[2]; }
        int tx, ty;
        if (!to_int_trunc(u + 0.5, tx) || !to_int_trunc(v + 0.5, ty)) continue;
        if (tx >= 0 && tx < P.cols && ty >= 0 && ty < P.rows)
            atomicMax(&winner[(size_t)ty * P.cols + tx], (unsigned int)(p + 1));
    }
}

// queues the vote: the quad kernel when the frame allows it (row-major depth, cols % 4 == 0, aligned pointers)
static int launch_vote(rsdsfm_ctx *ctx, const uint8_t *image, const double *depth, const double *R, const double *t, const SplatParams &P,
                       unsigned int *winner, float *coords3d)
{
    const long long total = (long long)P.rows * P.cols;
    const bool quads = P.layout != RSDSFM_DEPTH_COLMAJOR && P.cols % 4 == 0 && (reinterpret_cast<uintptr_t>(image) & 3) == 0 &&
                       (reinterpret_cast<uintptr_t>(depth) & 15) == 0;
    if (!quads) {
        k_splat_vote<<<grid_for(ctx, total, 8), kThreads, 0, ctx->stream>>>(image, depth, R, t, P, winner, coords3d);
        ctx->launches++;
        return RSDSFM_OK;
    }
    RS_TRY(ensure(ctx, ctx->splat_tab, sizeof(double) * (13 * (size_t)P.rows + (size_t)P.cols + 4)));
    double *tab = (double *)ctx->splat_tab.p;
    const int nmax = P.rows > P.cols ? P.rows : P.cols;
    k_splat_tables<<<(nmax + 127) / 128, 128, 0, ctx->stream>>>(R, t, P, tab);
    const dim3 grid((unsigned)((P.cols / 4 + kThreads - 1) / kThreads), (unsigned)P.rows);
    k_splat_vote4<<<grid, kThreads, 0, ctx->stream>>>(image, depth, R, t, tab, P, winner, coords3d);
    ctx->launches += 2;
    return RSDSFM_OK;
}

__global__ void __launch_bounds__(kThreads) k_splat_gather(const uint8_t *__restrict__ image,
                                                           const unsigned int *__restrict__ winner, long long total,
                                                           uint8_t *gs)
{
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        const unsigned int wv = winner[p];
        uint8_t b = 0, g = 0, r = 0;
        if (wv) { const size_t q = (size_t)(wv - 1) * 3; b = image[q]; g = image[q + 1]; r = image[q + 2]; }
        gs[3 * p] = b; gs[3 * p + 1] = g; gs[3 * p + 2] = r;
    }
}

// ---------------------------------------------------------------- a15: crack fill
__device__ __forceinline__ bool is_black(const uint8_t *p)
{   // cv::norm(Vec3b) <= 15  <=>  b^2+g^2+r^2 <= 225
    const int s = (int)p[0] * p[0] + (int)p[1] * p[1] + (int)p[2] * p[2];
    return s <= 225;
}
__device__ __forceinline__ uint8_t saturate_u8(double v)
{   // cv::saturate_cast<uchar>(double) = cvRound (half to even) + clamp
    const double r = rint(v);
    return (uint8_t)(r < 0.0 ? 0.0 : (r > 255.0 ? 255.0 : r));
}

__global__ void __launch_bounds__(kThreads) k_fill_cracks(const uint8_t *__restrict__ in, int rows, int cols, int off,
                                                          uint8_t *out)
{
    const long long total = (long long)rows * cols;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(p / cols), col = (int)(p - (long long)row * cols);
        const uint8_t *c = in + 3 * p;
        uint8_t o0 = c[0], o1 = c[1], o2 = c[2];
        if (row >= off && row < rows - off && col >= off && col < cols - off && is_black(c)) {
            const uint8_t *nb[4] = {in + 3 * (p - (long long)off * cols), in + 3 * (p + (long long)off * cols),
                                    in + 3 * (p - off), in + 3 * (p + off)};
            double s0 = 0, s1 = 0, s2 = 0;
            unsigned count = 0;
            for (int a = 0; a < 4; ++a)
                if (!is_black(nb[a])) { s0 += nb[a][0]; s1 += nb[a][1]; s2 += nb[a][2]; count++; }
            if (count > 0) {
                const double f = 1 / (double)count;
                o0 = saturate_u8(f * s0); o1 = saturate_u8(f * s1); o2 = saturate_u8(f * s2);
            }
        }
        out[3 * p] = o0; out[3 * p + 1] = o1; out[3 * p + 2] = o2;
    }
}

// Fused a14 gather + a15 crack fill (offset 1) for the drivers that do not hand the cracky GS image
// out: a CTA gathers a (kFH+2) x (kFW+2) tile of winning source colours into shared memory (one
// halo pixel all round) and fills from there; every thread produces 4 consecutive pixels = 12
// bytes, written as three 32-bit words when the row pitch allows it.
constexpr int kFW = 128, kFH = 8;
static_assert(kFW * kFH == 4 * kThreads, "4 pixels per thread");

__global__ void __launch_bounds__(kThreads) k_gather_fill(const uint8_t *__restrict__ image,
                                                          const unsigned int *__restrict__ winner, int rows, int cols,
                                                          uint8_t *__restrict__ out)
{
    __shared__ uchar4 tile[kFH + 2][kFW + 2];
    const int x0 = blockIdx.x * kFW, y0 = blockIdx.y * kFH;
    // two dependent gathers per tile entry (winner -> source pixel): issue all winner loads of this thread
    // before the first colour load so that their latencies overlap instead of adding up
    constexpr int kEntries = (kFH + 2) * (kFW + 2), kPer = (kEntries + kThreads - 1) / kThreads;
    unsigned int wv[kPer];
#pragma unroll
    for (int r = 0; r < kPer; ++r) {
        const int idx = threadIdx.x + r * kThreads;
        const int ty = idx / (kFW + 2), tx = idx - ty * (kFW + 2);
        const int y = y0 - 1 + ty, x = x0 - 1 + tx;
        wv[r] = (idx < kEntries && y >= 0 && y < rows && x >= 0 && x < cols) ? winner[(size_t)y * cols + x] : 0u;
    }
#pragma unroll
    for (int r = 0; r < kPer; ++r) {
        const int idx = threadIdx.x + r * kThreads;
        if (idx >= kEntries) break;
        const int ty = idx / (kFW + 2), tx = idx - ty * (kFW + 2);
        uchar4 px = make_uchar4(0, 0, 0, 0);
        if (wv[r]) { const size_t q = (size_t)(wv[r] - 1) * 3; px = make_uchar4(image[q], image[q + 1], image[q + 2], 0); }
        tile[ty][tx] = px;
    }
    __syncthreads();
    const int ly = threadIdx.x / (kFW / 4), lx = (threadIdx.x % (kFW / 4)) * 4;
    const int y = y0 + ly;
    if (y >= rows) return;
    uint8_t o[12];
    int nvalid = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int x = x0 + lx + j;
        const uchar4 c = tile[ly + 1][lx + j + 1];
        uint8_t o0 = c.x, o1 = c.y, o2 = c.z;
        const bool black = ((int)c.x * c.x + (int)c.y * c.y + (int)c.z * c.z) <= 225;      // cv::norm(Vec3b) <= 15
        if (x < cols) nvalid = j + 1;
        if (x < cols && y >= 1 && y < rows - 1 && x >= 1 && x < cols - 1 && black) {
            const uchar4 nb[4] = {tile[ly][lx + j + 1], tile[ly + 2][lx + j + 1], tile[ly + 1][lx + j], tile[ly + 1][lx + j + 2]};
            double s0 = 0, s1 = 0, s2 = 0;
            unsigned count = 0;
#pragma unroll
            for (int a = 0; a < 4; ++a)
                if (((int)nb[a].x * nb[a].x + (int)nb[a].y * nb[a].y + (int)nb[a].z * nb[a].z) > 225) {
                    s0 += nb[a].x; s1 += nb[a].y; s2 += nb[a].z; count++;
                }
            if (count > 0) {
                const double f = 1 / (double)count;
                o0 = saturate_u8(f * s0); o1 = saturate_u8(f * s1); o2 = saturate_u8(f * s2);
            }
        }
        o[3 * j] = o0; o[3 * j + 1] = o1; o[3 * j + 2] = o2;
    }
    const size_t base = 3 * ((size_t)y * cols + x0 + lx);
    if (nvalid == 4 && (reinterpret_cast<uintptr_t>(out + base) & 3) == 0) {       // the caller's pointer need not be word aligned
        unsigned int *w = reinterpret_cast<unsigned int *>(out + base);
        w[0] = o[0] | (o[1] << 8) | (o[2] << 16) | ((unsigned int)o[3] << 24);
        w[1] = o[4] | (o[5] << 8) | (o[6] << 16) | ((unsigned int)o[7] << 24);
        w[2] = o[8] | (o[9] << 8) | (o[10] << 16) | ((unsigned int)o[11] << 24);
    } else {
        for (int j = 0; j < 3 * nvalid; ++j) out[base + j] = o[j];
    }
}

// ---------------------------------------------------------------- host-side launchers (device pointers)
// zrows (nullable): nzrows x {sum z, max z, max -z} already produced by the kernel that wrote z (the LM
// solve's epilogue); otherwise one pass over z computes them.
int glue_device(rsdsfm_ctx *ctx, double *z, int zs, const double *xyz, int xs, int m, const double *K4, int rows,
                int cols, double z_min_init, int layout, double *depth_map, uint8_t *depth_img, double *stats /*device, 8*/,
                const double *zrows, int nzrows)
{
    const int grid = grid_for(ctx, m, 4);
    RS_TRY(ensure(ctx, ctx->rpart, sizeof(double) * 3 * (size_t)grid));
    RS_CUDA(ctx, cudaMemsetAsync(depth_map, 0, sizeof(double) * (size_t)rows * cols, ctx->stream));
    if (depth_img) RS_CUDA(ctx, cudaMemsetAsync(depth_img, 0, (size_t)rows * cols, ctx->stream));
    if (m > 0) {
        if (!zrows) {
            k_glue_reduce<<<grid, kThreads, 0, ctx->stream>>>(z, zs, m, (double *)ctx->rpart.p);
            ctx->launches++;
            zrows = (const double *)ctx->rpart.p;
            nzrows = grid;
        }
        k_glue_stats<<<1, 96, 0, ctx->stream>>>(zrows, nzrows, stats, m, z_min_init);
        ctx->launches++;
        k_glue_raster<<<grid, kThreads, 0, ctx->stream>>>(z, zs, xyz, xs, m, stats, K4[0], K4[1], K4[2], K4[3], rows,
                                                          cols, layout, depth_map, depth_img);
        ctx->launches++;
    }
    return RSDSFM_OK;
}

int poses_device(rsdsfm_ctx *ctx, const double *motion7_dev, const double *stats_dev, double gamma, int rows, double *R,
                 double *t)
{
    k_set_relative_pose<<<(rows + 127) / 128, 128, 0, ctx->stream>>>(motion7_dev, stats_dev, gamma, rows, R, t);
    ctx->launches++;
    return RSDSFM_OK;
}

int backproject_device(rsdsfm_ctx *ctx, const uint8_t *image, const double *depth, int layout, int rows, int cols,
                       const double *K4, const double *R, const double *t, int gs_mode, uint8_t *gs_out, float *coords3d)
{
    const long long total = (long long)rows * cols;
    RS_TRY(ensure(ctx, ctx->winner, sizeof(unsigned int) * (size_t)total));
    RS_CUDA(ctx, cudaMemsetAsync(ctx->winner.p, 0, sizeof(unsigned int) * (size_t)total, ctx->stream));
    SplatParams P{K4[0], K4[1], K4[2], K4[3], rows, cols, layout, gs_mode};
    const int grid = grid_for(ctx, total, 8);
    RS_TRY(launch_vote(ctx, image, depth, R, t, P, (unsigned int *)ctx->winner.p, coords3d));
    k_splat_gather<<<grid, kThreads, 0, ctx->stream>>>(image, (const unsigned int *)ctx->winner.p, total, gs_out);
    ctx->launches++;
    return RSDSFM_OK;
}

// a13/a14 + a15 with offset 1 in one go: vote, then gather + fill straight into `rectified`
int backproject_fill_device(rsdsfm_ctx *ctx, const uint8_t *image, const double *depth, int layout, int rows, int cols,
                            const double *K4, const double *R, const double *t, int gs_mode, uint8_t *rectified)
{
    const long long total = (long long)rows * cols;
    RS_TRY(ensure(ctx, ctx->winner, sizeof(unsigned int) * (size_t)total));
    RS_CUDA(ctx, cudaMemsetAsync(ctx->winner.p, 0, sizeof(unsigned int) * (size_t)total, ctx->stream));
    SplatParams P{K4[0], K4[1], K4[2], K4[3], rows, cols, layout, gs_mode};
    RS_TRY(launch_vote(ctx, image, depth, R, t, P, (unsigned int *)ctx->winner.p, nullptr));
    const dim3 grid((unsigned)((cols + kFW - 1) / kFW), (unsigned)((rows + kFH - 1) / kFH));
    k_gather_fill<<<grid, kThreads, 0, ctx->stream>>>(image, (const unsigned int *)ctx->winner.p, rows, cols, rectified);
    ctx->launches++;
    return RSDSFM_OK;
}

int fill_cracks_device(rsdsfm_ctx *ctx, const uint8_t *in, int rows, int cols, unsigned offset, uint8_t *out)
{
    const long long total = (long long)rows * cols;
    const int grid = grid_for(ctx, total, 8);
    k_fill_cracks<<<grid, kThreads, 0, ctx->stream>>>(in, rows, cols, (int)offset, out);
    ctx->launches++;
    return RSDSFM_OK;
}

// ---------------------------------------------------------------- reprojection error
// Camera::meanReprojectionError / createErrorImage (camera.cc:503-691) with
// RsFrame::getGroundtruthDepthMap (rsframe.cc:416-436).  Two per-pixel map-reduce passes:
//   k_reproj_points  ground-truth depth under the ORIGINAL scanline pose, ground-truth world point
//                    under the RELOCATED pose (float, like the reference's Vec3f), per-component scale
//                    est/true with the |s| > 10 outlier rule; sums: scale, valid entries, outliers
//   k_reproj_error   mean scale -> per-pixel Euclidean error, < 50 rule, optional 8-bit error image
// poses: rows x 24 doubles = original R[9], t[3], relocated R[9], t[3] per scanline.
struct ReprojParams {
    double fx, fy, cx, cy, max_norm;
    int rows, cols, layout;
};

__global__ void __launch_bounds__(kThreads) k_reproj_points(const float *__restrict__ est, const double *__restrict__ ux,
                                                            const double *__restrict__ uy, const double *__restrict__ uz,
                                                            const double *__restrict__ poses,
                                                            const double *__restrict__ depth_est, ReprojParams P,
                                                            float *__restrict__ truep, double *__restrict__ gt_depth,
                                                            double *__restrict__ partials)
{
    const long long total = (long long)P.rows * P.cols;
    double s[3] = {0.0, 0.0, 0.0}, mx[1] = {0.0};
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(p / P.cols), x = (int)(p - (long long)y * P.cols);
        const size_t mi = (P.layout == RSDSFM_DEPTH_COLMAJOR) ? ((size_t)y + (size_t)x * P.rows) : (size_t)p;
        const double *po = poses + 24 * (size_t)y;
        const double W[3] = {ux[mi], uy[mi], uz[mi]};
        double z = 0.0;
        if (sqrt(W[0] * W[0] + W[1] * W[1] + W[2] * W[2]) > 0)                       // rsframe.cc:428
            z = po[6] * W[0] + po[7] * W[1] + po[8] * W[2] + po[11] * 1.0;            // worldToCameraFrame(., y, false).z
        if (gt_depth) gt_depth[mi] = z;
        if (z == 0) z = depth_est[mi];                                               // planeToSpace default (rsframe.cc:657-659)
        const double nx = ((double)x - P.cx) * 1.0 / P.fx;
        const double ny = ((double)y - P.cy) * 1.0 / P.fy;
        const double Pc[3] = {z * nx, z * ny, z * 1.0};
        const double *Rs = po + 12, *ts = po + 21;                                   // relocated pose
        double Rt[9], ti[3];
        for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) Rt[a * 3 + c] = Rs[c * 3 + a];
        for (int a = 0; a < 3; ++a) ti[a] = (-Rt[a * 3 + 0]) * ts[0] + (-Rt[a * 3 + 1]) * ts[1] + (-Rt[a * 3 + 2]) * ts[2];
        for (int a = 0; a < 3; ++a) {
            const double Pw = Rt[a * 3 + 0] * Pc[0] + Rt[a * 3 + 1] * Pc[1] + Rt[a * 3 + 2] * Pc[2] + ti[a] * 1.0;
            const float pt = (float)Pw;
            truep[3 * p + a] = pt;
            const float sc = est[3 * p + a] / pt;                                    // camera.cc:632-634 (float division)
            if (fabsf(sc) > 10) s[2] += 1.0;                                         // outlier: entry zeroed, not averaged
            else if (sc != 0 && sc == sc) { s[0] += (double)sc; s[1] += 1.0; }
        }
    }
    block_reduce_store<3, 0>(s, mx, partials);
}

__global__ void __launch_bounds__(kThreads) k_reproj_error(const float *__restrict__ est, const float *__restrict__ truep,
                                                           const double *__restrict__ sums, ReprojParams P,
                                                           uint8_t *__restrict__ error_image, double *__restrict__ partials)
{
    const long long total = (long long)P.rows * P.cols;
    const double scale = sums[0] / sums[1];                                          // camera.cc:657
    double s[2] = {0.0, 0.0}, mx[1] = {0.0};
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        const double e0 = est[3 * p] / scale, e1 = est[3 * p + 1] / scale, e2 = est[3 * p + 2] / scale;
        const double t0 = truep[3 * p], t1 = truep[3 * p + 1], t2 = truep[3 * p + 2];
        const double d0 = e0 - t0, d1 = e1 - t1, d2 = e2 - t2;
        const double nrm = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
        if (e0 == e0 && e1 == e1 && e2 == e2 && t0 == t0 && t1 == t1 && t2 == t2 && nrm < 50) { s[0] += nrm; s[1] += 1.0; }
        if (error_image) {                                                           // createErrorImage, camera.cc:583
            int q;
            if (!to_int_trunc(nrm * 255 / P.max_norm + 0.5, q)) q = (int)0x80000000;
            error_image[p] = (uint8_t)(q & 0xff);
        }
    }
    block_reduce_store<2, 0>(s, mx, partials);
}

// sums5 (device, 8 doubles): [0] sum of scales, [1] entries averaged, [2] outliers, [3] sum of errors, [4] points used
int reproj_device(rsdsfm_ctx *ctx, const float *est, const double *ux, const double *uy, const double *uz,
                  const double *poses24, const double *depth_est, int layout, int rows, int cols, const double *K4,
                  double max_norm, float *truep, uint8_t *error_image, double *gt_depth, double *sums5)
{
    const long long total = (long long)rows * cols;
    const int grid = grid_for(ctx, total, 4);
    RS_TRY(ensure(ctx, ctx->rpart, sizeof(double) * 3 * (size_t)grid));
    double *partials = (double *)ctx->rpart.p;
    ReprojParams P{K4[0], K4[1], K4[2], K4[3], max_norm, rows, cols, layout};
    k_reproj_points<<<grid, kThreads, 0, ctx->stream>>>(est, ux, uy, uz, poses24, depth_est, P, truep, gt_depth, partials);
    ctx->launches++;
    launch_final_reduce(ctx, partials, grid, 3, 0, sums5);
    k_reproj_error<<<grid, kThreads, 0, ctx->stream>>>(est, truep, sums5, P, error_image, partials);
    ctx->launches++;
    launch_final_reduce(ctx, partials, grid, 2, 0, sums5 + 3);
    RS_CUDA(ctx, cudaGetLastError());
    return RSDSFM_OK;
}

// ---------------------------------------------------------------- ground-truth flow
// Camera::calculateTrueFlow (camera.cc:209-249) + RsFrame::calculateImageCoordinatesRsFrame
// (rsframe.cc:740-768): for every pixel of frame 1 its world point is projected with EVERY scanline
// pose of frame 2 and the pose whose row index is closest to the projected y wins (first minimum).
// O(rows^2 cols) projections -- one thread per pixel, the poses stream through shared memory in
// chunks that every thread of the CTA walks in the same order (broadcast reads).  FP64-issue bound:
// one IEEE division per (pixel, scanline); plain IEEE sequence => bit-exact against the oracle.
constexpr int kPoseChunk = 128;            // scanline poses per shared-memory chunk (12 doubles each)

__global__ void __launch_bounds__(kThreads) k_true_flow(const double *__restrict__ ux, const double *__restrict__ uy,
                                                        const double *__restrict__ uz, const double *__restrict__ poses2,
                                                        ReprojParams P, double2 *__restrict__ flow)
{
    __shared__ double sp[kPoseChunk * 12];
    const long long total = (long long)P.rows * P.cols;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = p < total;
    const int v = in ? (int)(p / P.cols) : 0, u = in ? (int)(p - (long long)v * P.cols) : 0;
    const size_t mi = (P.layout == RSDSFM_DEPTH_COLMAJOR) ? ((size_t)v + (size_t)u * P.rows) : (size_t)(in ? p : 0);
    const double W0 = in ? ux[mi] : 0.0, W1 = in ? uy[mi] : 0.0, W2 = in ? uz[mi] : 0.0;
    const bool solid = in && sqrt(W0 * W0 + W1 * W1 + W2 * W2) != 0;                 // camera.cc:230
    double min_diff = INFINITY;
    int best = 0;
    for (int c0 = 0; c0 < P.rows; c0 += kPoseChunk) {
        const int nc = (P.rows - c0 < kPoseChunk) ? (P.rows - c0) : kPoseChunk;
        __syncthreads();
        for (int j = threadIdx.x; j < nc * 12; j += blockDim.x) sp[j] = poses2[(size_t)c0 * 12 + j];
        __syncthreads();
        if (solid)
            for (int i = 0; i < nc; ++i) {
                const double *R = sp + 12 * i;
                const double Y = R[3] * W0 + R[4] * W1 + R[5] * W2 + R[10] * 1.0;
                const double Z = R[6] * W0 + R[7] * W1 + R[8] * W2 + R[11] * 1.0;
                const double qy = Y / Z * P.fx + P.cy;                               // spaceToPlane: f_x for y too (Q12)
                const double diff = fabs(qy - (double)(c0 + i));
                if (diff < min_diff) { min_diff = diff; best = c0 + i; }
            }
    }
    if (!in) return;
    double px = (double)u, py = (double)v;
    if (solid) {
        const double *R = poses2 + 12 * (size_t)best;
        const double X = R[0] * W0 + R[1] * W1 + R[2] * W2 + R[9] * 1.0;
        const double Y = R[3] * W0 + R[4] * W1 + R[5] * W2 + R[10] * 1.0;
        const double Z = R[6] * W0 + R[7] * W1 + R[8] * W2 + R[11] * 1.0;
        const double bx = X / Z * P.fx + P.cx, by = Y / Z * P.fx + P.cy;
        if (sqrt(bx * bx + by * by) != 0) { px = bx; py = by; }                      // camera.cc:236-238
    }
    flow[p] = make_double2(px - (double)u, py - (double)v);
}

// poses2: device, rows x 12 doubles (R[9] row-major, t[3]) of frame 2's scanlines
int true_flow_device(rsdsfm_ctx *ctx, const double *ux, const double *uy, const double *uz, const double *poses2, int layout,
                     int rows, int cols, const double *K4, double *flow)
{
    const long long total = (long long)rows * cols;
    ReprojParams P{K4[0], K4[1], K4[2], K4[3], 1.0, rows, cols, layout};
    k_true_flow<<<(unsigned)((total + kThreads - 1) / kThreads), kThreads, 0, ctx->stream>>>(ux, uy, uz, poses2, P, (double2 *)flow);
    ctx->launches++;
    RS_CUDA(ctx, cudaGetLastError());
    return RSDSFM_OK;
}

}  // namespace rsdsfm
