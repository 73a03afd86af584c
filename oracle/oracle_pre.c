/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the reference's preprocess stage.
 *
 * The reference (aruco_detect.py) delegates all arithmetic to OpenCV (un-vendored third-party dependency,
 * reference README.md:42: opencv-contrib 4.2.0; the runnable copy in this image is opencv-python-headless
 * 4.13.0).  This file restates the published algorithms of the OpenCV calls on the path, in plain C, and is
 * pinned against the cv2 4.13 binary by tests/test_oracle_*.py (the reference itself ships no golden
 * vectors for this path: "parity unpinned" by the reference's own tests, pinned here against its
 * dependency's outputs).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Call sites restated:
 *   aruco_detect.py:568  cv2.initUndistortRectifyMap(mtx, dist, None, mtx, (w,h), CV_32FC1)
 *   aruco_detect.py:252  cv2.remap(frame, mapx, mapy, INTER_LINEAR)          (BORDER_CONSTANT 0)
 *   aruco_detect.py:255  cv2.cvtColor(frame, COLOR_RGB2LAB)                  (8-bit integer path)
 *   aruco_detect.py:256  cv2.LUT(lab[...,0], lookUpTable)
 *   aruco_detect.py:257  cv2.cvtColor(lab, COLOR_LAB2RGB)                    (8-bit integer path)
 *   aruco_detect.py:592  cv2.cvtColor(frame, COLOR_BGR2GRAY)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------------ */
/* aruco_detect.py:568 -- rational distortion model, R = I, newK = K.  dist: k1 k2 p1 p2 k3 k4 k5 k6 (s,tau = 0) */
void orc_init_undistort_map(const double *K, const double *D, int w, int h, float *mapx, float *mapy)
{
    double fx = K[0], fy = K[4], u0 = K[2], v0 = K[5];
    double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3], k3 = D[4], k4 = D[5], k5 = D[6], k6 = D[7];
    double s1 = D[8], s2 = D[9], s3 = D[10], s4 = D[11];
    /* inverse of the (upper-triangular, zero-skew) new camera matrix */
    double ir0 = 1.0 / fx, ir2 = -u0 / fx, ir4 = 1.0 / fy, ir5 = -v0 / fy;
    for (int i = 0; i < h; i++) {
        for (int j = 0; j < w; j++) {
            double x = j * ir0 + ir2, y = i * ir4 + ir5;
            double x2 = x * x, y2 = y * y;
            double r2 = x2 + y2, _2xy = 2 * x * y;
            double kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((k6 * r2 + k5) * r2 + k4) * r2);
            double xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2) + s1 * r2 + s2 * r2 * r2;
            double yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy + s3 * r2 + s4 * r2 * r2;
            mapx[(size_t)i * w + j] = (float)(fx * xd + u0);
            mapy[(size_t)i * w + j] = (float)(fy * yd + v0);
        }
    }
}

/* ------------------------------------------------------------------------------------------------------ */
/* aruco_detect.py:252 -- bilinear remap, Q5 coordinates, Q15 weights */
static int16_t g_wtab[32 * 32][4];
static int g_wtab_ready = 0;

static void build_wtab(void)
{
    for (int fy = 0; fy < 32; fy++)
        for (int fx = 0; fx < 32; fx++) {
            float ay = fy * (1.f / 32), ax = fx * (1.f / 32);
            float wy[2] = {1.f - ay, ay}, wx[2] = {1.f - ax, ax};
            for (int k = 0; k < 4; k++) {
                float v = wy[k >> 1] * wx[k & 1] * 32768.f;
                long r = lrintf(v);
                if (r > 32767) r = 32767;
                g_wtab[fy * 32 + fx][k] = (int16_t)r;
            }
        }
    g_wtab_ready = 1;
}

void orc_remap_bilinear(const uint8_t *src, int sw, int sh, int cn, const float *mapx, const float *mapy,
                        int dw, int dh, uint8_t *dst)
{
    if (!g_wtab_ready) build_wtab();
    for (int y = 0; y < dh; y++)
        for (int x = 0; x < dw; x++) {
            size_t o = (size_t)y * dw + x;
            /* float32 product then round-half-even (saturate_cast<int>(float)) */
            int sx = (int)lrintf(mapx[o] * 32.f), sy = (int)lrintf(mapy[o] * 32.f);
            int ix = sx >> 5, iy = sy >> 5;
            const int16_t *wt = g_wtab[(sy & 31) * 32 + (sx & 31)];
            for (int c = 0; c < cn; c++) {
                int acc = 0;
                for (int k = 0; k < 4; k++) {
                    int xx = ix + (k & 1), yy = iy + (k >> 1);
                    int p = (xx >= 0 && xx < sw && yy >= 0 && yy < sh) ? src[((size_t)yy * sw + xx) * cn + c] : 0;
                    acc += wt[k] * p;
                }
                dst[o * cn + c] = (uint8_t)((acc + 16384) >> 15);
            }
        }
}

/* ------------------------------------------------------------------------------------------------------ */
/* aruco_detect.py:255-257 -- 8-bit sRGB <-> CIELab (D65) integer pipeline */
static uint16_t t_gamma[256];      /* sRGB gamma, scale 2040          */
static uint16_t t_cbrt[3072];      /* cube-root curve, scale 32768    */
static uint16_t t_ly[256], t_lf[256];
static uint8_t t_invgamma[4096];
static int g_lab_ready = 0;

static const int FWD[9] = {1777, 1541, 778, 871, 2929, 296, 73, 448, 3575};
static const int INV[9] = {12615, -6296, -2223, -3773, 7684, 185, 217, -836, 4715};
#define LAB_BASE 16384

static void build_lab_tables(void)
{
    for (int i = 0; i < 256; i++) {
        float x = i * (1.f / 255.f);
        double g = x <= 0.04045f ? (double)(x * (1.f / 12.92f)) : (double)(float)pow((double)(x + 0.055) * (1. / 1.055), 2.4);
        t_gamma[i] = (uint16_t)lrint(2040.0 * g);
    }
    for (int i = 0; i < 3072; i++) {
        float x = i * (1.f / 2040.f);
        double c = x < 0.008856f ? (double)(x * 7.787f + 0.13793103448275862f) : cbrt((double)x);
        t_cbrt[i] = (uint16_t)lrint(32768.0 * c);
    }
    /* The dependency evaluates the cube root with a float32 rational approximation; over the reachable
       index range it differs from exact rounding in exactly two entries (pinned by the exhaustive 2^24
       colour test against cv2 4.13: tests/test_oracle_pre.py). */
    t_cbrt[49] = 9454;
    t_cbrt[628] = 22126;
    for (int i = 0; i < 256; i++) {
        int y, f;
        if (i <= 20) {
            float yy = (float)(i * LAB_BASE * 100) * 27.f / (float)(255 * 24389);
            y = (int)lrintf(yy);
            float ff = (float)LAB_BASE * (16.f / 116.f + (float)(i * 100 * 841) * 27.f / (float)(255 * 24389) / 108.f);
            f = (int)lrintf(ff);
        } else {
            float fy = ((float)(i * 100) / 255.f + 16.f) / 116.f;
            f = (int)lrintf((float)LAB_BASE * fy);
            y = (int)lrintf((float)LAB_BASE * fy * fy * fy);
        }
        t_ly[i] = (uint16_t)y;
        t_lf[i] = (uint16_t)f;
    }
    for (int i = 0; i < 4096; i++) {
        float x = i * (1.f / 4096.f);
        double v = x <= 0.0031308 ? x * 12.92 : 1.055 * pow((double)x, 1. / 2.4) - 0.055;
        long r = lrint(255.0 * v);
        t_invgamma[i] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
    }
    g_lab_ready = 1;
}

static inline int clip255(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }

void orc_rgb2lab_u8(const uint8_t *src, size_t npx, uint8_t *dst)
{
    if (!g_lab_ready) build_lab_tables();
    for (size_t i = 0; i < npx; i++) {
        int R = t_gamma[src[3 * i]], G = t_gamma[src[3 * i + 1]], B = t_gamma[src[3 * i + 2]];
        int fX = t_cbrt[(R * FWD[0] + G * FWD[1] + B * FWD[2] + 2048) >> 12];
        int fY = t_cbrt[(R * FWD[3] + G * FWD[4] + B * FWD[5] + 2048) >> 12];
        int fZ = t_cbrt[(R * FWD[6] + G * FWD[7] + B * FWD[8] + 2048) >> 12];
        dst[3 * i] = (uint8_t)clip255((296 * fY - 1336934 + 16384) >> 15);
        dst[3 * i + 1] = (uint8_t)clip255((500 * (fX - fY) + 128 * 32768 + 16384) >> 15);
        dst[3 * i + 2] = (uint8_t)clip255((200 * (fY - fZ) + 128 * 32768 + 16384) >> 15);
    }
}

static inline int ab_to_xz(int v)
{
    if (v <= 3390) return v * 108 / 841 - LAB_BASE * 16 / 116 * 108 / 841;
    return v * v / LAB_BASE * v / LAB_BASE;
}

void orc_lab2rgb_u8(const uint8_t *src, size_t npx, uint8_t *dst)
{
    if (!g_lab_ready) build_lab_tables();
    for (size_t i = 0; i < npx; i++) {
        int L = src[3 * i], a = src[3 * i + 1], b = src[3 * i + 2];
        int y = t_ly[L], f = t_lf[L];
        int adiv = ((5 * a * 53687 + 128) >> 13) - 128 * LAB_BASE / 500;
        int bdiv = ((b * 41943 + 16) >> 9) - 128 * LAB_BASE / 200 + 1;
        int X = ab_to_xz(f + adiv), Z = ab_to_xz(f - bdiv);
        for (int c = 0; c < 3; c++) {
            int v = (INV[3 * c] * X + INV[3 * c + 1] * y + INV[3 * c + 2] * Z + 8192) >> 14;
            v = v < 0 ? 0 : v > 4095 ? 4095 : v;
            dst[3 * i + c] = t_invgamma[v];
        }
    }
}

void orc_lut_channel0(uint8_t *lab, size_t npx, const uint8_t *lut)
{
    for (size_t i = 0; i < npx; i++) lab[3 * i] = lut[lab[3 * i]];
}

/* aruco_detect.py:592 */
void orc_bgr2gray(const uint8_t *src, size_t npx, uint8_t *dst)
{
    for (size_t i = 0; i < npx; i++)
        dst[i] = (uint8_t)((src[3 * i] * 3735 + src[3 * i + 1] * 19235 + src[3 * i + 2] * 9798 + 16384) >> 15);
}

/* aruco_detect.py:250-259 + :592 in one call (remap -> Lab gamma -> gray) */
void orc_preprocess(const uint8_t *bgr, int w, int h, const float *mapx, const float *mapy, const uint8_t *lut,
                    uint8_t *bgr_out, uint8_t *gray)
{
    size_t n = (size_t)w * h;
    uint8_t *tmp = (uint8_t *)malloc(n * 3);
    orc_remap_bilinear(bgr, w, h, 3, mapx, mapy, w, h, bgr_out);
    orc_rgb2lab_u8(bgr_out, n, tmp);
    orc_lut_channel0(tmp, n, lut);
    orc_lab2rgb_u8(tmp, n, bgr_out);
    orc_bgr2gray(bgr_out, n, gray);
    free(tmp);
}

/* expose tables so that tests can diff them */
void orc_lab_tables(uint16_t *gamma, uint16_t *cbrt_, uint16_t *ly, uint16_t *lf, uint8_t *invgamma)
{
    if (!g_lab_ready) build_lab_tables();
    memcpy(gamma, t_gamma, sizeof t_gamma);
    memcpy(cbrt_, t_cbrt, sizeof t_cbrt);
    memcpy(ly, t_ly, sizeof t_ly);
    memcpy(lf, t_lf, sizeof t_lf);
    memcpy(invgamma, t_invgamma, sizeof t_invgamma);
}

/* test hook: override one cube-root table entry (used once to pin the dependency's float rounding) */
void orc_set_cbrt_entry(int i, int v)
{
    if (!g_lab_ready) build_lab_tables();
    t_cbrt[i] = (uint16_t)v;
}

/* ------------------------------------------------------------------------------------------------------ */
/* cv2.undistort(src, K, D[, newK]) (dcnn/scripts/tests/visualize_uav.py:62): the dependency builds
 * CV_16SC2 maps, i.e. the FP64 source coordinate is scaled by 32 and rounded directly (no float32 map in between),
 * then the same Q5 / Q15 bilinear remap with BORDER_CONSTANT 0 */
void orc_undistort(const uint8_t *src, int w, int h, int cn, const double *K, const double *D, const double *newK, uint8_t *dst)
{
    if (!g_wtab_ready) build_wtab();
    const double *A = newK ? newK : K;
    double fx = K[0], fy = K[4], u0 = K[2], v0 = K[5];
    double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3], k3 = D[4], k4 = D[5], k5 = D[6], k6 = D[7];
    double s1 = D[8], s2 = D[9], s3 = D[10], s4 = D[11];
    double ir0 = 1.0 / A[0], ir2 = -A[2] / A[0], ir4 = 1.0 / A[4], ir5 = -A[5] / A[4];
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            double x = j * ir0 + ir2, y = i * ir4 + ir5;
            double x2 = x * x, y2 = y * y;
            double r2 = x2 + y2, _2xy = 2 * x * y;
            double kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((k6 * r2 + k5) * r2 + k4) * r2);
            double xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2) + s1 * r2 + s2 * r2 * r2;
            double yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy + s3 * r2 + s4 * r2 * r2;
            double u = fx * xd + u0, v = fy * yd + v0;
            long sx = lrint(u * 32.0), sy = lrint(v * 32.0);
            if (sx < -2147483647L) sx = -2147483647L;
            if (sx > 2147483647L) sx = 2147483647L;
            if (sy < -2147483647L) sy = -2147483647L;
            if (sy > 2147483647L) sy = 2147483647L;
            int ix = (int)(sx >> 5), iy = (int)(sy >> 5);
            const int16_t *wt = g_wtab[(sy & 31) * 32 + (sx & 31)];
            size_t o = (size_t)i * w + j;
            for (int c = 0; c < cn; c++) {
                int acc = 0;
                for (int k = 0; k < 4; k++) {
                    int xx = ix + (k & 1), yy = iy + (k >> 1);
                    int p = (xx >= 0 && xx < w && yy >= 0 && yy < h) ? src[((size_t)yy * w + xx) * cn + c] : 0;
                    acc += wt[k] * p;
                }
                dst[o * cn + c] = (uint8_t)((acc + 16384) >> 15);
            }
        }
}

/* ---------------------------------------------------------------------------------------------------------
 * The image the APRILTAG quad detector works on when aprilTagQuadDecimate / aprilTagQuadSigma are set (dependency: aruco
 * detectMarkers, APRILTAG branch; knobs documented at aruco_detect.py:203,231-233): cv2.resize(INTER_AREA) by 1/f for an
 * integer factor f, then cv2.GaussianBlur (sigma > 0) or unsharp masking (sigma < 0) with floor(4 |sigma|) | 1 taps.
 * Pinned against cv2.resize / cv2.GaussianBlur in tests/test_oracle_pre.py. */
void orc_resize_area_int(const uint8_t *src, int w, int h, int f, int dw, int dh, uint8_t *dst)
{
    const float scale = 1.f / (float)(f * f);
    for (int y = 0; y < dh; y++)
        for (int x = 0; x < dw; x++) {
            int s = 0, cnt = 0;
            for (int dy = 0; dy < f && y * f + dy < h; dy++)
                for (int dx = 0; dx < f && x * f + dx < w; dx++) { s += src[(size_t)(y * f + dy) * w + x * f + dx]; cnt++; }
            int v;
            if (cnt == f * f) v = f == 2 ? (s + 2) >> 2 : (int)lrintf((float)s * scale);
            else v = cnt ? (int)lrintf((float)s / (float)cnt) : 0;   /* partial block at the right / bottom edge: mean of what is there */
            dst[(size_t)y * dw + x] = (uint8_t)v;
        }
}

/* cv2.resize(INTER_AREA) for a non-integer shrink factor (the reference's own example is aprilTagQuadDecimate = 1.5,
 * aruco_detect.py:203): the dependency's table-driven area filter in float32.  Tap table of one axis: for every destination
 * index the source pixels it covers with weights coverage / cell width; a destination pixel is
 * sum_rows beta * (sum_cols alpha * src), accumulated in float32 in table order, then rounded half to even. */
typedef struct { int si, di; float alpha; } orc_area_tap;

int orc_area_tab(int ssize, int dsize, double scale, orc_area_tap *tab)
{
    int k = 0;
    for (int dx = 0; dx < dsize; dx++) {
        double fsx1 = dx * scale, fsx2 = fsx1 + scale;
        double cell = scale < ssize - fsx1 ? scale : ssize - fsx1;
        int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
        if (sx2 > ssize - 1) sx2 = ssize - 1;
        if (sx1 > sx2) sx1 = sx2;
        if (sx1 - fsx1 > 1e-3) { tab[k].di = dx; tab[k].si = sx1 - 1; tab[k++].alpha = (float)((sx1 - fsx1) / cell); }
        for (int sx = sx1; sx < sx2; sx++) { tab[k].di = dx; tab[k].si = sx; tab[k++].alpha = (float)(1.0 / cell); }
        if (fsx2 - sx2 > 1e-3) {
            double a = fsx2 - sx2;
            if (a > 1.) a = 1.;
            if (a > cell) a = cell;
            tab[k].di = dx; tab[k].si = sx2; tab[k++].alpha = (float)(a / cell);
        }
    }
    return k;
}

/* destination size of cv2.resize(src, None, fx = 1 / decimate (float32 division), fy = ...): cvRound of the product */
void orc_resize_area_dsize(int w, int h, float decimate, int *dw, int *dh)
{
    const double inv = (double)(1.f / decimate);
    *dw = (int)lrint(w * inv); *dh = (int)lrint(h * inv);
}

/* scale: the dependency keeps 1 / fx (fx = 1.f / decimate in float32) as the filter scale even where the rounded destination
 * size would imply a slightly different one */
void orc_resize_area(const uint8_t *src, int w, int h, int dw, int dh, double scale, uint8_t *dst)
{
    const double sx = scale, sy = scale;
    orc_area_tap *xt = (orc_area_tap *)malloc(sizeof(orc_area_tap) * (size_t)w * 2), *yt = (orc_area_tap *)malloc(sizeof(orc_area_tap) * (size_t)h * 2);
    const int nx = orc_area_tab(w, dw, sx, xt), ny = orc_area_tab(h, dh, sy, yt);
    float *buf = (float *)malloc(sizeof(float) * dw), *sum = (float *)malloc(sizeof(float) * dw);
    int prev_dy = ny > 0 ? yt[0].di : 0;
    for (int x = 0; x < dw; x++) sum[x] = 0;
    for (int j = 0; j < ny; j++) {
        const float beta = yt[j].alpha;
        const int dy = yt[j].di;
        const uint8_t *S = src + (size_t)yt[j].si * w;
        for (int x = 0; x < dw; x++) buf[x] = 0;
        for (int k = 0; k < nx; k++) buf[xt[k].di] += S[xt[k].si] * xt[k].alpha;
        if (dy != prev_dy) {
            for (int x = 0; x < dw; x++) {
                long v = lrintf(sum[x]);
                dst[(size_t)prev_dy * dw + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
                sum[x] = beta * buf[x];
            }
            prev_dy = dy;
        } else {
            for (int x = 0; x < dw; x++) sum[x] += beta * buf[x];
        }
    }
    for (int x = 0; x < dw; x++) {
        long v = lrintf(sum[x]);
        dst[(size_t)prev_dy * dw + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
    }
    free(buf); free(sum); free(xt); free(yt);
}

/* 8.8 fixed-point Gaussian kernel with error diffusion towards the centre tap; returns the tap count (0: too many) */
int orc_gauss_kernel_fixed(float sigma_f, int *k, int max_taps)
{
    const float s = fabsf(sigma_f);
    int ksz = (int)floorf(4 * s);
    ksz |= 1;
    if (ksz <= 1) { k[0] = 256; return 1; }
    if (ksz > max_taps) return 0;
    const double sigma = (double)s, scale2x = -0.5 * 0.25 / (sigma * sigma);
    const int n2 = (ksz - 1) / 2;
    double vals[64], sum = 0;
    for (int i = 0, x = 1 - ksz; i < n2; i++, x += 2) { vals[i] = exp((double)(x * x) * scale2x); sum += vals[i]; }
    sum = sum * 2 + 1.0;
    const double mul1 = 1.0 / sum;
    double err = 0;
    long long tot = 0;
    for (int i = 0; i < n2; i++) {
        const double adj = vals[i] * mul1 * 256.0 + err;
        const long long v0 = llrint(adj);
        err = adj - (double)v0;
        k[i] = k[ksz - 1 - i] = (int)v0;
        tot += v0;
    }
    k[n2] = (int)(256 - 2 * tot);
    return ksz;
}

/* sigma > 0: blur; sigma < 0: clamp(2 * src - blur); |sigma| below 0.5: copy.  Returns 0, or -1 when the kernel is too long */
int orc_quad_sigma(const uint8_t *src, int w, int h, float sigma, uint8_t *dst)
{
    int k[64];
    const int n = orc_gauss_kernel_fixed(sigma, k, 63);
    if (n == 0) return -1;
    if (n == 1 || sigma == 0) { memcpy(dst, src, (size_t)w * h); return 0; }
    const int r = n / 2;
    uint16_t *tmp = (uint16_t *)malloc((size_t)w * h * sizeof(uint16_t));
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int s = 0;
            for (int i = 0; i < n; i++) { int xx = x + i - r; xx = xx < 0 ? 0 : xx >= w ? w - 1 : xx; s += k[i] * src[(size_t)y * w + xx]; }
            tmp[(size_t)y * w + x] = (uint16_t)s;
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            unsigned s = 0;
            for (int i = 0; i < n; i++) { int yy = y + i - r; yy = yy < 0 ? 0 : yy >= h ? h - 1 : yy; s += (unsigned)k[i] * tmp[(size_t)yy * w + x]; }
            int v = (int)((s + 32768u) >> 16);
            if (sigma < 0) { v = 2 * src[(size_t)y * w + x] - v; v = v < 0 ? 0 : v > 255 ? 255 : v; }
            dst[(size_t)y * w + x] = (uint8_t)v;
        }
    free(tmp);
    return 0;
}
