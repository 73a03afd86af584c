// preprocess.cu -- stage 1 of the marker pipeline on B200 (sm_100a):
//   K0  build_undistort_map   aruco_detect.py:568  cv2.initUndistortRectifyMap
//   K1  preprocess_fused      aruco_detect.py:250-259 + :592  remap + RGB2LAB + LUT(L) + LAB2RGB + BGR2GRAY
//   stand-alone remap / cvtColor / LUT kernels for the drop-in cv2-shaped API
// Integer pipeline, bit-exact with the dependency's 8-bit paths.  HBM-bound by design:
// one pass over the BGR frame (24.9 MB read) producing gray (8.3 MB) and optionally the corrected BGR frame.
#include "common.cuh"
#include <math.h>
#include <string.h>

// ---------------------------------------------------------------------------------------------------------
// host: integer colour tables (values are what OpenCV's RGB2Lab_b / Lab2RGBinteger tables hold)
static void build_lab_tables_host(LabTables &T, const uint8_t *lut /* nullable = identity */)
{
    for (int i = 0; i < 256; i++) {
        float x = i * (1.f / 255.f);
        double g = x <= 0.04045f ? (double)(x * (1.f / 12.92f)) : (double)(float)pow((double)(x + 0.055) * (1. / 1.055), 2.4);
        T.gamma[i] = (uint16_t)lrint(2040.0 * g);
    }
    for (int i = 0; i < 3072; i++) {
        float x = i * (1.f / 2040.f);
        double c = x < 0.008856f ? (double)(x * 7.787f + 0.13793103448275862f) : cbrt((double)x);
        T.cbrt[i] = (uint16_t)lrint(32768.0 * c);
    }
    // float32 cube-root approximation of the dependency rounds these two entries the other way
    T.cbrt[49] = 9454;
    T.cbrt[628] = 22126;
    uint16_t ly[256], lf[256];
    const int BASE = 16384;
    for (int i = 0; i < 256; i++) {
        int y, f;
        if (i <= 20) {
            y = (int)lrintf((float)(i * BASE * 100) * 27.f / (float)(255 * 24389));
            f = (int)lrintf((float)BASE * (16.f / 116.f + (float)(i * 100 * 841) * 27.f / (float)(255 * 24389) / 108.f));
        } else {
            float fy = ((float)(i * 100) / 255.f + 16.f) / 116.f;
            f = (int)lrintf((float)BASE * fy);
            y = (int)lrintf((float)BASE * fy * fy * fy);
        }
        ly[i] = (uint16_t)y;
        lf[i] = (uint16_t)f;
    }
    for (int i = 0; i < 256; i++) {  // compose with the gamma LUT on L (aruco_detect.py:256)
        int j = lut ? lut[i] : i;
        T.ly[i] = ly[j];
        T.lf[i] = lf[j];
    }
    for (int i = 0; i < 4096; i++) {
        float x = i * (1.f / 4096.f);
        double v = x <= 0.0031308 ? x * 12.92 : 1.055 * pow((double)x, 1. / 2.4) - 0.055;
        long r = lrint(255.0 * v);
        T.invgamma[i] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
    }
}

int apse_upload_tables(apse_ctx *ctx, const uint8_t *lut, LabTables **dev, cudaStream_t st)
{
    LabTables T;
    build_lab_tables_host(T, lut);
    if (!*dev) CUDA_TRY(ctx, cudaMalloc((void **)dev, sizeof(LabTables)));
    CUDA_TRY(ctx, cudaMemcpyAsync(*dev, &T, sizeof(LabTables), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));  // T is a stack object
    return APSE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// device: colour chain on one pixel (channel 0 is treated as "R" exactly as COLOR_RGB2LAB does on BGR data)
struct SmemTables {
    uint16_t gamma[256];
    uint16_t cbrt[3072];
    uint16_t ly[256];
    uint16_t lf[256];
    uint8_t invgamma[4096];
};

__device__ __forceinline__ void load_tables(SmemTables *s, const LabTables *g)
{
    const uint32_t *src = reinterpret_cast<const uint32_t *>(g);
    uint32_t *dst = reinterpret_cast<uint32_t *>(s);
    for (int i = threadIdx.x; i < (int)(sizeof(SmemTables) / 4); i += blockDim.x) dst[i] = src[i];
}

__device__ __forceinline__ int clip255(int v) { return min(max(v, 0), 255); }

__device__ __forceinline__ void rgb2lab_px(const SmemTables *T, int c0, int c1, int c2, int &L, int &a, int &b)
{
    int R = T->gamma[c0], G = T->gamma[c1], B = T->gamma[c2];
    int fX = T->cbrt[(R * 1777 + G * 1541 + B * 778 + 2048) >> 12];
    int fY = T->cbrt[(R * 871 + G * 2929 + B * 296 + 2048) >> 12];
    int fZ = T->cbrt[(R * 73 + G * 448 + B * 3575 + 2048) >> 12];
    L = clip255((296 * fY - 1336934 + 16384) >> 15);
    a = clip255((500 * (fX - fY) + 128 * 32768 + 16384) >> 15);
    b = clip255((200 * (fY - fZ) + 128 * 32768 + 16384) >> 15);
}

__device__ __forceinline__ int ab_to_xz(int v)
{
    // integer divisions truncate toward zero, as in the dependency's table construction
    return v <= 3390 ? v * 108 / 841 - 290 : v * v / 16384 * v / 16384;
}

__device__ __forceinline__ void lab2rgb_px(const SmemTables *T, int L, int a, int b, int &o0, int &o1, int &o2)
{
    int y = T->ly[L], f = T->lf[L];
    int adiv = ((5 * a * 53687 + 128) >> 13) - 4194;
    int bdiv = ((b * 41943 + 16) >> 9) - 10485 + 1;
    int X = ab_to_xz(f + adiv), Z = ab_to_xz(f - bdiv);
    int r0 = (12615 * X - 6296 * y - 2223 * Z + 8192) >> 14;
    int r1 = (-3773 * X + 7684 * y + 185 * Z + 8192) >> 14;
    int r2 = (217 * X - 836 * y + 4715 * Z + 8192) >> 14;
    o0 = T->invgamma[min(max(r0, 0), 4095)];
    o1 = T->invgamma[min(max(r1, 0), 4095)];
    o2 = T->invgamma[min(max(r2, 0), 4095)];
}

__device__ __forceinline__ int gray_px(int c0, int c1, int c2) { return (c0 * 3735 + c1 * 19235 + c2 * 9798 + 16384) >> 15; }

// Q5 fixed-point source coordinate of the dependency's remap: rint(map * 32) evaluated in float32
__device__ __forceinline__ int q5(float m) { return __float2int_rn(__fmul_rn(m, 32.f)); }

// ---------------------------------------------------------------------------------------------------------
// K0: undistort map (FP64 per pixel, init only)
__global__ void k_build_undistort_map(int w, int h, double fx, double fy, double u0, double v0, double k1, double k2,
                                      double p1, double p2, double k3, double k4, double k5, double k6, double s1,
                                      double s2, double s3, double s4, float *__restrict__ mapx,
                                      float *__restrict__ mapy)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= w || i >= h) return;
    double ir0 = 1.0 / fx, ir2 = -u0 / fx, ir4 = 1.0 / fy, ir5 = -v0 / fy;
    double x = __dadd_rn(__dmul_rn(j, ir0), ir2), y = __dadd_rn(__dmul_rn(i, ir4), ir5);
    double x2 = __dmul_rn(x, x), y2 = __dmul_rn(y, y);
    double r2 = __dadd_rn(x2, y2), _2xy = __dmul_rn(__dmul_rn(2, x), y);
    double num = __dadd_rn(1, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k3, r2), k2), r2), k1), r2));
    double den = __dadd_rn(1, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k6, r2), k5), r2), k4), r2));
    double kr = __ddiv_rn(num, den);
    double xd = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, kr), __dmul_rn(p1, _2xy)),
                                              __dmul_rn(p2, __dadd_rn(r2, __dmul_rn(2, x2)))),
                                    __dmul_rn(s1, r2)),
                          __dmul_rn(__dmul_rn(s2, r2), r2));
    double yd = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(y, kr), __dmul_rn(p1, __dadd_rn(r2, __dmul_rn(2, y2)))),
                                              __dmul_rn(p2, _2xy)),
                                    __dmul_rn(s3, r2)),
                          __dmul_rn(__dmul_rn(s4, r2), r2));
    size_t o = (size_t)i * w + j;
    mapx[o] = (float)__dadd_rn(__dmul_rn(fx, xd), u0);
    mapy[o] = (float)__dadd_rn(__dmul_rn(fy, yd), v0);
}

int apse_init_undistort_map(apse_ctx *ctx, const double K[9], const double D[14], int w, int h, float *mapx,
                            float *mapy, void *stream)
{
    if (!ctx || !K || !D || !mapx || !mapy || w <= 0 || h <= 0) CTX_FAIL(ctx, APSE_ERR_INVALID_ARG, "init_undistort_map: bad argument");
    if (D[12] != 0 || D[13] != 0) CTX_FAIL(ctx, APSE_ERR_UNSUPPORTED, "tilted sensor model (tauX/tauY) is not supported");
    dim3 block(256), grid(div_up(w, 256), h);
    KLAUNCH(ctx, KID_BUILD_MAP, (cudaStream_t)stream, k_build_undistort_map<<<grid, block, 0, (cudaStream_t)stream>>>(w, h, K[0], K[4], K[2], K[5], D[0], D[1], D[2], D[3],
                                                                    D[4], D[5], D[6], D[7], D[8], D[9], D[10], D[11],
                                                                    mapx, mapy));
    return APSE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// bilinear tap set of one output pixel
struct Taps {
    int off00;      // byte offset of tap (ix,iy) channel 0 in the source frame
    int w00, w01, w10, w11;
    unsigned mask;  // bit k set = tap k inside the image (k = 0:(ix,iy) 1:(ix+1,iy) 2:(ix,iy+1) 3:(ix+1,iy+1))
};

__device__ __forceinline__ Taps make_taps(float mx, float my, int sw, int sh, int cn)
{
    Taps t;
    int sx = q5(mx), sy = q5(my);
    int ix = sx >> 5, iy = sy >> 5, fx = sx & 31, fy = sy & 31;
    t.w00 = min(32767, (32 - fy) * (32 - fx) * 32);
    t.w01 = (32 - fy) * fx * 32;
    t.w10 = fy * (32 - fx) * 32;
    t.w11 = fy * fx * 32;
    bool x0 = ix >= 0 && ix < sw, x1 = ix + 1 >= 0 && ix + 1 < sw;
    bool y0 = iy >= 0 && iy < sh, y1 = iy + 1 >= 0 && iy + 1 < sh;
    t.mask = (x0 && y0 ? 1u : 0u) | (x1 && y0 ? 2u : 0u) | (x0 && y1 ? 4u : 0u) | (x1 && y1 ? 8u : 0u);
    t.off00 = (iy * sw + ix) * cn;
    return t;
}

__device__ __forceinline__ int sample(const uint8_t *__restrict__ src, const Taps &t, int rowbytes, int cn, int c)
{
    const uint8_t *p = src + t.off00 + c;
    int acc = 16384;
    if (t.mask == 15u) {
        acc += t.w00 * __ldg(p) + t.w01 * __ldg(p + cn) + t.w10 * __ldg(p + rowbytes) + t.w11 * __ldg(p + rowbytes + cn);
    } else {
        if (t.mask & 1u) acc += t.w00 * __ldg(p);
        if (t.mask & 2u) acc += t.w01 * __ldg(p + cn);
        if (t.mask & 4u) acc += t.w10 * __ldg(p + rowbytes);
        if (t.mask & 8u) acc += t.w11 * __ldg(p + rowbytes + cn);
    }
    return acc >> 15;
}

// K1: fused preprocess.  Block = 256 threads = 8 rows x 32 lanes, each lane 4 consecutive pixels (128 x 8 tile).
// The map of the tile is read once and kept in registers while the block walks `fpb` frames of the batch, so
// map traffic is amortised over the batch; gray is written as one 32-bit word per lane (128 B per warp).
#define K1_PX 4
__global__ void __launch_bounds__(256, 4) k_preprocess_fused(const uint8_t *__restrict__ bgr, uint8_t *__restrict__ bgr_out,
                                                          uint8_t *__restrict__ gray, const float *__restrict__ mapx,
                                                          const float *__restrict__ mapy, const LabTables *__restrict__ tables,
                                                          int w, int h, int batch, int fpb)
{
    __shared__ SmemTables T;
    load_tables(&T, tables);
    __syncthreads();
    int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
    int x0 = (blockIdx.x * 32 + lane) * K1_PX, y = blockIdx.y * 8 + row;
    if (y >= h || x0 >= w) return;
    Taps taps[K1_PX];
    int npx = min(K1_PX, w - x0);
#pragma unroll
    for (int k = 0; k < K1_PX; k++) {
        int xx = min(x0 + k, w - 1);
        size_t o = (size_t)y * w + xx;
        taps[k] = make_taps(__ldg(mapx + o), __ldg(mapy + o), w, h, 3);
    }
    size_t frame_px = (size_t)w * h;
    int f0 = blockIdx.z * fpb, f1 = min(batch, f0 + fpb);
    for (int f = f0; f < f1; f++) {
        const uint8_t *src = bgr + (size_t)f * frame_px * 3;
        int g[K1_PX], o0[K1_PX], o1[K1_PX], o2[K1_PX];
#pragma unroll
        for (int k = 0; k < K1_PX; k++) {
            int c0 = sample(src, taps[k], w * 3, 3, 0);
            int c1 = sample(src, taps[k], w * 3, 3, 1);
            int c2 = sample(src, taps[k], w * 3, 3, 2);
            int L, a, b;
            rgb2lab_px(&T, c0, c1, c2, L, a, b);
            lab2rgb_px(&T, L, a, b, o0[k], o1[k], o2[k]);
            g[k] = gray_px(o0[k], o1[k], o2[k]);
        }
        size_t o = (size_t)f * frame_px + (size_t)y * w + x0;
        if (npx == K1_PX && (w & 3) == 0) {
            *reinterpret_cast<uint32_t *>(gray + o) = (uint32_t)g[0] | ((uint32_t)g[1] << 8) | ((uint32_t)g[2] << 16) | ((uint32_t)g[3] << 24);
            if (bgr_out) {
                uint32_t *d = reinterpret_cast<uint32_t *>(bgr_out + o * 3);
                d[0] = (uint32_t)o0[0] | ((uint32_t)o1[0] << 8) | ((uint32_t)o2[0] << 16) | ((uint32_t)o0[1] << 24);
                d[1] = (uint32_t)o1[1] | ((uint32_t)o2[1] << 8) | ((uint32_t)o0[2] << 16) | ((uint32_t)o1[2] << 24);
                d[2] = (uint32_t)o2[2] | ((uint32_t)o0[3] << 8) | ((uint32_t)o1[3] << 16) | ((uint32_t)o2[3] << 24);
            }
        } else {
            for (int k = 0; k < npx; k++) {
                gray[o + k] = (uint8_t)g[k];
                if (bgr_out) {
                    bgr_out[(o + k) * 3] = (uint8_t)o0[k];
                    bgr_out[(o + k) * 3 + 1] = (uint8_t)o1[k];
                    bgr_out[(o + k) * 3 + 2] = (uint8_t)o2[k];
                }
            }
        }
    }
}

int apse_preprocess(apse_ctx *ctx, const uint8_t *bgr, uint8_t *bgr_out, uint8_t *gray, int batch, void *stream)
{
    if (!ctx || !bgr || !gray || batch <= 0) CTX_FAIL(ctx, APSE_ERR_INVALID_ARG, "preprocess: bad argument");
    if (!ctx->has_camera || !ctx->has_lut) CTX_FAIL(ctx, APSE_ERR_NOT_CONFIGURED, "preprocess: set_camera and set_lut first");
    int w = ctx->w, h = ctx->h;
    int fpb = batch >= 8 ? 8 : batch;
    dim3 grid(div_up(w, 32 * K1_PX), div_up(h, 8), div_up(batch, fpb));
    KLAUNCH(ctx, KID_PREPROCESS, (cudaStream_t)stream, k_preprocess_fused<<<grid, 256, 0, (cudaStream_t)stream>>>(bgr, bgr_out, gray, ctx->mapx, ctx->mapy, ctx->tables, w, h,
                                                              batch, fpb));
    return APSE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// stand-alone kernels for the cv2-shaped drop-in calls
__global__ void k_remap(const uint8_t *__restrict__ src, int sw, int sh, int cn, const float *__restrict__ mapx,
                        const float *__restrict__ mapy, int dw, int dh, uint8_t *__restrict__ dst)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw || y >= dh) return;
    size_t o = (size_t)y * dw + x;
    Taps t = make_taps(__ldg(mapx + o), __ldg(mapy + o), sw, sh, cn);
    for (int c = 0; c < cn; c++) dst[o * cn + c] = (uint8_t)sample(src, t, sw * cn, cn, c);
}

int apse_remap(apse_ctx *ctx, const uint8_t *src, int sw, int sh, int cn, const float *mapx, const float *mapy, int dw,
               int dh, uint8_t *dst, void *stream)
{
    if (!ctx || !src || !mapx || !mapy || !dst || sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0)
        CTX_FAIL(ctx, APSE_ERR_INVALID_ARG, "remap: bad argument");
    if (cn != 1 && cn != 3) CTX_FAIL(ctx, APSE_ERR_UNSUPPORTED, "remap: only 1- or 3-channel 8-bit images");
    if ((int64_t)sw * sh * cn >= (1ll << 31)) CTX_FAIL(ctx, APSE_ERR_UNSUPPORTED, "remap: source image too large");
    dim3 grid(div_up(dw, 256), dh);
    KLAUNCH(ctx, KID_REMAP, (cudaStream_t)stream, k_remap<<<grid, 256, 0, (cudaStream_t)stream>>>(src, sw, sh, cn, mapx, mapy, dw, dh, dst));
    return APSE_OK;
}

__global__ void __launch_bounds__(256) k_rgb2lab(const uint8_t *__restrict__ src, int64_t npx, uint8_t *__restrict__ dst,
                                                 const LabTables *__restrict__ tables)
{
    __shared__ SmemTables T;
    load_tables(&T, tables);
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (int64_t)gridDim.x * blockDim.x) {
        int L, a, b;
        rgb2lab_px(&T, src[3 * i], src[3 * i + 1], src[3 * i + 2], L, a, b);
        dst[3 * i] = (uint8_t)L;
        dst[3 * i + 1] = (uint8_t)a;
        dst[3 * i + 2] = (uint8_t)b;
    }
}

__global__ void __launch_bounds__(256) k_lab2rgb(const uint8_t *__restrict__ src, int64_t npx, uint8_t *__restrict__ dst,
                                                 const LabTables *__restrict__ tables)
{
    __shared__ SmemTables T;
    load_tables(&T, tables);
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (int64_t)gridDim.x * blockDim.x) {
        int o0, o1, o2;
        lab2rgb_px(&T, src[3 * i], src[3 * i + 1], src[3 * i + 2], o0, o1, o2);
        dst[3 * i] = (uint8_t)o0;
        dst[3 * i + 1] = (uint8_t)o1;
        dst[3 * i + 2] = (uint8_t)o2;
    }
}

__global__ void k_bgr2gray(const uint8_t *__restrict__ src, int64_t npx, uint8_t *__restrict__ dst)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = (uint8_t)gray_px(src[3 * i], src[3 * i + 1], src[3 * i + 2]);
}

__global__ void k_lut(const uint8_t *__restrict__ src, int64_t n, int sstride, const uint8_t *__restrict__ lut,
                      uint8_t *__restrict__ dst, int dstride)
{
    __shared__ uint8_t L[256];
    if (threadIdx.x < 256) L[threadIdx.x] = lut[threadIdx.x];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i * dstride] = L[src[i * sstride]];
}

static int grid_for(int64_t n) { return (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16); }

int apse_cvt_rgb2lab(apse_ctx *ctx, const uint8_t *src, int64_t npx, uint8_t *dst, void *stream)
{
    if (!ctx || !src || !dst || npx <= 0) CTX_FAIL(ctx, APSE_ERR_INVALID_ARG, "cvt_rgb2lab: bad argument");
    KLAUNCH(ctx, KID_CVT, (cudaStream_t)stream, k_rgb2lab<<<grid_for(npx), 256, 0, (cudaStream_t)stream>>>(src, npx, dst, ctx->tables_id));
    return APSE_OK;
}
int apse_cvt_lab2rgb(apse_ctx *ctx, const uint8_t *src, int64_t npx, uint8_t *dst, void *stream)
{
    if (!ctx || !src || !dst || npx <= 0) CTX_FAIL(ctx, APSE_ERR_INVALID_ARG, "cvt_lab2rgb: bad argument");
    KLAUNCH(ctx, KID_CVT, (cudaStream_t)stream, k_lab2rgb<<<grid_for(npx), 256, 0, (cudaStream_t)stream>>>(src, npx, dst, ctx->tables_id));
    return APSE_OK;
}
int apse_cvt_bgr2gray(apse_ctx *ctx, const uint8_t *src, int64_t npx, uint8_t *dst, void *stream)
{
    if (!ctx || !src || !dst || npx <= 0) CTX_FAIL(ctx, APSE_ERR_INVALID_ARG, "cvt_bgr2gray: bad argument");
    KLAUNCH(ctx, KID_CVT, (cudaStream_t)stream, k_bgr2gray<<<grid_for(npx), 256, 0, (cudaStream_t)stream>>>(src, npx, dst));
    return APSE_OK;
}
int apse_lut(apse_ctx *ctx, const uint8_t *src, int64_t n, int src_stride, const uint8_t *lut_dev, uint8_t *dst,
             int dst_stride, void *stream)
{
    if (!ctx || !src || !dst || !lut_dev || n <= 0 || src_stride <= 0 || dst_stride <= 0)
        CTX_FAIL(ctx, APSE_ERR_INVALID_ARG, "lut: bad argument");
    KLAUNCH(ctx, KID_LUT, (cudaStream_t)stream, k_lut<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(src, n, src_stride, lut_dev, dst, dst_stride));
    return APSE_OK;
}
