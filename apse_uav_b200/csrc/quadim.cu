// quadim.cu -- the image the APRILTAG quad detector actually looks at when aprilTagQuadDecimate / aprilTagQuadSigma are set
// (aruco_detect.py:203,231-233 document the knobs; the reference leaves them off).  The dependency (aruco detectMarkers, APRILTAG
// mode) shrinks the gray frame by 1/decimate with INTER_AREA, blurs it (sigma > 0) or sharpens it by unsharp masking (sigma < 0)
// with a (floor(4 |sigma|) | 1)-tap Gaussian, finds the quads on that image and scales their corners back by `decimate`;
// identification then samples the ORIGINAL frame.  Arithmetic restated from the dependency's 8-bit paths and pinned against
// cv2.resize / cv2.GaussianBlur by the CPU restatement in tests/ (test_*_pre.py):
//   INTER_AREA, integer factor f: block sum, f == 2: (s + 2) >> 2, else rint(float(s) * (1.f / (f * f))) (float32, half to even),
//   partial blocks at the right / bottom edge: rint(float(s) / count); any other factor (the reference's example is 1.5): the
//   dependency's table-driven area filter, sum_rows beta * (sum_cols alpha * src) in float32 in table order, scale = 1 / fx with
//   fx = 1.f / decimate
//   GaussianBlur 8-bit: 8.8 fixed-point kernel with error diffusion towards the centre tap (sum exactly 256), horizontal pass
//   exact in 16 bits, vertical pass in 32 bits, (v + 32768) >> 16, BORDER_REPLICATE.
#include "common.cuh"
#include <math.h>
#include <string.h>
#include <vector>

#define QI_MAX_TAPS 33

struct GaussKernel { int n; int k[QI_MAX_TAPS]; };

// host: the dependency's bit-exact fixed-point Gaussian kernel (8 fractional bits)
static bool make_gauss_kernel(float sigma_f, GaussKernel &K)
{
    const float s = fabsf(sigma_f);
    int ksz = (int)floorf(4 * s);
    ksz |= 1;
    if (ksz <= 1) { K.n = 1; K.k[0] = 256; return true; }
    if (ksz > QI_MAX_TAPS) return false;
    const double sigma = (double)s, scale2x = -0.5 * 0.25 / (sigma * sigma);
    const int n2 = (ksz - 1) / 2;
    double vals[QI_MAX_TAPS], sum = 0;
    for (int i = 0, x = 1 - ksz; i < n2; i++, x += 2) { vals[i] = exp((double)(x * x) * scale2x); sum += vals[i]; }
    sum = sum * 2 + 1.0;
    const double mul1 = 1.0 / sum;
    double err = 0;
    long long tot = 0;
    for (int i = 0; i < n2; i++) {
        const double adj = vals[i] * mul1 * 256.0 + err;
        const long long v0 = llrint(adj);
        err = adj - (double)v0;
        K.k[i] = K.k[ksz - 1 - i] = (int)v0;
        tot += v0;
    }
    K.k[n2] = (int)(256 - 2 * tot);
    K.n = ksz;
    return true;
}

__global__ void k_decimate(const uint8_t *__restrict__ src, int w, int h, int f, float scale, uint8_t *__restrict__ dst, int dw, int dh)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, fr = blockIdx.z;
    if (x >= dw || y >= dh) return;
    const uint8_t *p = src + (size_t)fr * w * h + (size_t)y * f * w + x * f;
    int s = 0, cnt = 0;
    for (int dy = 0; dy < f && y * f + dy < h; dy++)
        for (int dx = 0; dx < f && x * f + dx < w; dx++) { s += p[(size_t)dy * w + dx]; cnt++; }
    int v;
    if (cnt == f * f) v = f == 2 ? (s + 2) >> 2 : __float2int_rn(__fmul_rn((float)s, scale));
    else v = cnt ? __float2int_rn(__fdiv_rn((float)s, (float)cnt)) : 0;
    dst[(size_t)fr * dw * dh + (size_t)y * dw + x] = (uint8_t)v;
}

// non-integer factor: taps of destination column x = xt[xo[x] .. xo[x + 1]), of destination row y = yt[yo[y] .. yo[y + 1])
struct AreaTap { int si; float alpha; };
__global__ void k_resize_area(const uint8_t *__restrict__ src, int w, int h, const int *__restrict__ xo, const AreaTap *__restrict__ xt,
                              const int *__restrict__ yo, const AreaTap *__restrict__ yt, uint8_t *__restrict__ dst, int dw, int dh)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, fr = blockIdx.z;
    if (x >= dw || y >= dh) return;
    const uint8_t *im = src + (size_t)fr * w * h;
    const int x0 = xo[x], x1 = xo[x + 1], y0 = yo[y], y1 = yo[y + 1];
    float sum = 0.f;
    for (int j = y0; j < y1; j++) {
        const uint8_t *S = im + (size_t)yt[j].si * w;
        float buf = 0.f;
        for (int k = x0; k < x1; k++) buf = __fadd_rn(buf, __fmul_rn((float)S[xt[k].si], xt[k].alpha));
        sum = __fadd_rn(sum, __fmul_rn(yt[j].alpha, buf));
    }
    dst[(size_t)fr * dw * dh + (size_t)y * dw + x] = (uint8_t)min(max(__float2int_rn(sum), 0), 255);
}

// host: the dependency's tap table of one axis (computeResizeAreaTab)
static void area_tab(int ssize, int dsize, double scale, std::vector<int> &ofs, std::vector<AreaTap> &tab)
{
    ofs.assign(dsize + 1, 0);
    tab.clear();
    for (int dx = 0; dx < dsize; dx++) {
        ofs[dx] = (int)tab.size();
        const double fsx1 = dx * scale, fsx2 = fsx1 + scale, cell = scale < ssize - fsx1 ? scale : ssize - fsx1;
        int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
        if (sx2 > ssize - 1) sx2 = ssize - 1;
        if (sx1 > sx2) sx1 = sx2;
        if (sx1 - fsx1 > 1e-3) tab.push_back(AreaTap{sx1 - 1, (float)((sx1 - fsx1) / cell)});
        for (int sx = sx1; sx < sx2; sx++) tab.push_back(AreaTap{sx, (float)(1.0 / cell)});
        if (fsx2 - sx2 > 1e-3) {
            double a = fsx2 - sx2;
            if (a > 1.) a = 1.;
            if (a > cell) a = cell;
            tab.push_back(AreaTap{sx2, (float)(a / cell)});
        }
    }
    ofs[dsize] = (int)tab.size();
}

__global__ void k_gauss_h(const uint8_t *__restrict__ src, int w, int h, GaussKernel K, uint16_t *__restrict__ tmp)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, fr = blockIdx.z;
    if (x >= w) return;
    const uint8_t *row = src + (size_t)fr * w * h + (size_t)y * w;
    const int r = K.n >> 1;
    int s = 0;
    for (int i = 0; i < K.n; i++) s += K.k[i] * (int)row[min(max(x + i - r, 0), w - 1)];
    tmp[(size_t)fr * w * h + (size_t)y * w + x] = (uint16_t)s;
}

// SHARPEN: dst = clamp(2 * orig - blur) (sigma < 0), else dst = blur
__global__ void k_gauss_v(const uint16_t *__restrict__ tmp, int w, int h, GaussKernel K, const uint8_t *__restrict__ orig, int sharpen,
                          uint8_t *__restrict__ dst)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, fr = blockIdx.z;
    if (x >= w) return;
    const uint16_t *col = tmp + (size_t)fr * w * h + x;
    const int r = K.n >> 1;
    unsigned s = 0;
    for (int i = 0; i < K.n; i++) s += (unsigned)K.k[i] * (unsigned)col[(size_t)min(max(y + i - r, 0), h - 1) * w];
    int v = (int)((s + 32768u) >> 16);
    const size_t o = (size_t)fr * w * h + (size_t)y * w + x;
    if (sharpen) v = min(max(2 * (int)orig[o] - v, 0), 255);
    dst[o] = (uint8_t)v;
}

__global__ void k_scale_quads(float *__restrict__ quads, const int32_t *__restrict__ counters, float f)
{
    const int fr = blockIdx.y;
    const int nq = min(counters[fr * APSE_COUNTERS + 2], APSE_MAX_QUADS);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nq * 8; i += gridDim.x * blockDim.x) {
        float *q = quads + (size_t)fr * APSE_MAX_QUADS * 8 + i;
        *q = __fmul_rn(*q, f);
    }
}

// true when the quad detector must run on a derived image
bool apse_quad_image_needed(const apse_params &p)
{
    GaussKernel K;
    const bool blur = p.aprilTagQuadSigma != 0 && make_gauss_kernel(p.aprilTagQuadSigma, K) && K.n > 1;
    return p.aprilTagQuadDecimate > 1 || blur;
}

// gray [batch][h][w] -> ctx->quad_im [batch][qh][qw]; returns the detector's image size and the corner scale
int apse_quad_image(apse_ctx *ctx, const uint8_t *gray, int w, int h, int batch, const uint8_t **out, int *qw, int *qh, float *scale,
                    cudaStream_t st)
{
    const apse_params &p = ctx->params;
    const bool dec = p.aprilTagQuadDecimate > 1;
    // cv2.resize(src, None, fx = fy = 1.f / decimate): destination size = cvRound(size * fx), filter scale = 1 / fx
    const double fx = (double)(1.f / (dec ? p.aprilTagQuadDecimate : 1.f)), fscale = 1.0 / fx;
    const int dw = dec ? (int)lrint(w * fx) : w, dh = dec ? (int)lrint(h * fx) : h;
    const int f = (int)lrint(fscale);
    const bool integer_factor = fabs(fscale - f) < 2.220446049250313e-16;
    if (dec && (dw < 8 || dh < 8)) CTX_FAIL(ctx, APSE_ERR_INVALID_ARG, "detect: aprilTagQuadDecimate %.3f leaves a %dx%d image", p.aprilTagQuadDecimate, dw, dh);
    GaussKernel K;
    if (!make_gauss_kernel(p.aprilTagQuadSigma, K)) CTX_FAIL(ctx, APSE_ERR_UNSUPPORTED, "detect: aprilTagQuadSigma %.3f needs more than %d taps", p.aprilTagQuadSigma, QI_MAX_TAPS);
    const bool blur = p.aprilTagQuadSigma != 0 && K.n > 1;
    const size_t npx = (size_t)ctx->max_batch * ctx->max_w * ctx->max_h;
    if (!ctx->quad_im) {
        CUDA_TRY(ctx, cudaMalloc((void **)&ctx->quad_im, npx));
        CUDA_TRY(ctx, cudaMalloc((void **)&ctx->quad_im2, npx));
        CUDA_TRY(ctx, cudaMalloc((void **)&ctx->quad_tmp, npx * sizeof(uint16_t)));
    }
    if (dec && !integer_factor && (ctx->area_w != w || ctx->area_h != h || ctx->area_dec != p.aprilTagQuadDecimate)) {
        // tap tables of the area filter for this geometry (rebuilt when the frame size or the factor changes)
        std::vector<int> xo, yo;
        std::vector<AreaTap> xt, yt;
        area_tab(w, dw, fscale, xo, xt);
        area_tab(h, dh, fscale, yo, yt);
        const size_t bytes = (xo.size() + yo.size()) * sizeof(int) + (xt.size() + yt.size()) * sizeof(AreaTap);
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        cudaFree(ctx->area_tab);
        ctx->area_tab = nullptr;
        CUDA_TRY(ctx, cudaMalloc(&ctx->area_tab, bytes));
        std::vector<unsigned char> blob(bytes);
        size_t o = 0;
        ctx->area_ofs[0] = o; memcpy(blob.data() + o, xo.data(), xo.size() * sizeof(int)); o += xo.size() * sizeof(int);
        ctx->area_ofs[1] = o; memcpy(blob.data() + o, yo.data(), yo.size() * sizeof(int)); o += yo.size() * sizeof(int);
        ctx->area_ofs[2] = o; memcpy(blob.data() + o, xt.data(), xt.size() * sizeof(AreaTap)); o += xt.size() * sizeof(AreaTap);
        ctx->area_ofs[3] = o; memcpy(blob.data() + o, yt.data(), yt.size() * sizeof(AreaTap));
        CUDA_TRY(ctx, cudaMemcpy(ctx->area_tab, blob.data(), bytes, cudaMemcpyHostToDevice));
        ctx->area_w = w; ctx->area_h = h; ctx->area_dec = p.aprilTagQuadDecimate;
    }
    const uint8_t *cur = gray;
    if (dec && integer_factor) {
        KLAUNCH(ctx, KID_TILE_MINMAX, st, k_decimate<<<dim3(div_up(dw, 256), dh, batch), 256, 0, st>>>(gray, w, h, f, 1.f / (float)(f * f), ctx->quad_im, dw, dh));
        cur = ctx->quad_im;
    } else if (dec) {
        const unsigned char *t = (const unsigned char *)ctx->area_tab;
        KLAUNCH(ctx, KID_TILE_MINMAX, st, k_resize_area<<<dim3(div_up(dw, 256), dh, batch), 256, 0, st>>>(
                    gray, w, h, (const int *)(t + ctx->area_ofs[0]), (const AreaTap *)(t + ctx->area_ofs[2]), (const int *)(t + ctx->area_ofs[1]),
                    (const AreaTap *)(t + ctx->area_ofs[3]), ctx->quad_im, dw, dh));
        cur = ctx->quad_im;
    }
    if (blur) {
        uint8_t *dst = cur == ctx->quad_im ? ctx->quad_im2 : ctx->quad_im;
        KLAUNCH(ctx, KID_TILE_MINMAX, st, k_gauss_h<<<dim3(div_up(dw, 256), dh, batch), 256, 0, st>>>(cur, dw, dh, K, ctx->quad_tmp));
        KLAUNCH(ctx, KID_TILE_MINMAX, st, k_gauss_v<<<dim3(div_up(dw, 256), dh, batch), 256, 0, st>>>(ctx->quad_tmp, dw, dh, K, cur, p.aprilTagQuadSigma < 0 ? 1 : 0, dst));
        cur = dst;
    }
    *out = cur; *qw = dw; *qh = dh; *scale = dec ? p.aprilTagQuadDecimate : 1.f;
    return APSE_OK;
}

int apse_scale_quads(apse_ctx *ctx, int batch, float f, cudaStream_t st)
{
    KLAUNCH(ctx, KID_TILE_MINMAX, st, k_scale_quads<<<dim3(4, batch), 256, 0, st>>>(ctx->quads, ctx->counters, f));
    return APSE_OK;
}
