// quadim.cu -- the image the APRILTAG quad detector actually looks at when aprilTagQuadDecimate / aprilTagQuadSigma are set
// (aruco_detect.py:203,231-233 document the knobs; the reference leaves them off).  The dependency (aruco detectMarkers, APRILTAG
// mode) shrinks the gray frame by 1/decimate with INTER_AREA, blurs it (sigma > 0) or sharpens it by unsharp masking (sigma < 0)
// with a (floor(4 |sigma|) | 1)-tap Gaussian, finds the quads on that image and scales their corners back by `decimate`;
// identification then samples the ORIGINAL frame.  Arithmetic restated from the dependency's 8-bit paths and pinned against
// cv2.resize / cv2.GaussianBlur (oracle/oracle_pre.c, tests/test_oracle_pre.py):
//   INTER_AREA, integer factor f: block sum, f == 2: (s + 2) >> 2, else rint(float(s) * (1.f / (f * f))) (float32, half to even)
//   GaussianBlur 8-bit: 8.8 fixed-point kernel with error diffusion towards the centre tap (sum exactly 256), horizontal pass
//   exact in 16 bits, vertical pass in 32 bits, (v + 32768) >> 16, BORDER_REPLICATE.
#include "common.cuh"
#include <math.h>

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
    int s = 0;
    for (int dy = 0; dy < f; dy++)
        for (int dx = 0; dx < f; dx++) s += p[(size_t)dy * w + dx];
    const int v = f == 2 ? (s + 2) >> 2 : __float2int_rn(__fmul_rn((float)s, scale));
    dst[(size_t)fr * dw * dh + (size_t)y * dw + x] = (uint8_t)v;
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
    const int f = p.aprilTagQuadDecimate > 1 ? (int)p.aprilTagQuadDecimate : 1;
    if (p.aprilTagQuadDecimate > 1 && ((float)f != p.aprilTagQuadDecimate || w % f || h % f))
        CTX_FAIL(ctx, APSE_ERR_UNSUPPORTED, "detect: aprilTagQuadDecimate %.3f needs an integer factor that divides the %dx%d frame", p.aprilTagQuadDecimate, w, h);
    GaussKernel K;
    if (!make_gauss_kernel(p.aprilTagQuadSigma, K)) CTX_FAIL(ctx, APSE_ERR_UNSUPPORTED, "detect: aprilTagQuadSigma %.3f needs more than %d taps", p.aprilTagQuadSigma, QI_MAX_TAPS);
    const bool blur = p.aprilTagQuadSigma != 0 && K.n > 1;
    const size_t npx = (size_t)ctx->max_batch * ctx->max_w * ctx->max_h;
    if (!ctx->quad_im) {
        CUDA_TRY(ctx, cudaMalloc((void **)&ctx->quad_im, npx));
        CUDA_TRY(ctx, cudaMalloc((void **)&ctx->quad_im2, npx));
        CUDA_TRY(ctx, cudaMalloc((void **)&ctx->quad_tmp, npx * sizeof(uint16_t)));
    }
    const int dw = w / f, dh = h / f;
    const uint8_t *cur = gray;
    if (f > 1) {
        KLAUNCH(ctx, KID_TILE_MINMAX, st, k_decimate<<<dim3(div_up(dw, 256), dh, batch), 256, 0, st>>>(gray, w, h, f, 1.f / (float)(f * f), ctx->quad_im, dw, dh));
        cur = ctx->quad_im;
    }
    if (blur) {
        uint8_t *dst = cur == ctx->quad_im ? ctx->quad_im2 : ctx->quad_im;
        KLAUNCH(ctx, KID_TILE_MINMAX, st, k_gauss_h<<<dim3(div_up(dw, 256), dh, batch), 256, 0, st>>>(cur, dw, dh, K, ctx->quad_tmp));
        KLAUNCH(ctx, KID_TILE_MINMAX, st, k_gauss_v<<<dim3(div_up(dw, 256), dh, batch), 256, 0, st>>>(ctx->quad_tmp, dw, dh, K, cur, p.aprilTagQuadSigma < 0 ? 1 : 0, dst));
        cur = dst;
    }
    *out = cur; *qw = dw; *qh = dh; *scale = p.aprilTagQuadDecimate > 1 ? p.aprilTagQuadDecimate : 1.f;
    return APSE_OK;
}

int apse_scale_quads(apse_ctx *ctx, int batch, float f, cudaStream_t st)
{
    KLAUNCH(ctx, KID_TILE_MINMAX, st, k_scale_quads<<<dim3(4, batch), 256, 0, st>>>(ctx->quads, ctx->counters, f));
    return APSE_OK;
}
