// decode.cu -- candidate post-filter, perspective bit sampling and dictionary matching of
// aruco.detectMarkers (aruco_detect.py:267, dictionary aruco_detect.py:263); SURVEY.md rows a6.A5-a6.A7,
// cv2 4.13 semantics.  One CTA per frame:
//   * quads ordered deterministically, border filter, stable sort by perimeter (descending)
//   * "too close" predicate for all pairs in parallel (bit matrix), sequential grouping walk on one thread
//   * every candidate decoded by one warp: FP64 homography, 48x48 nearest-neighbour gather, Otsu on a
//     shared-memory histogram, per-cell majority, border check, Hamming match against the dictionary
//   * nesting hierarchy + depth-ordered acceptance, corner rotation, output
// Compiled with -fmad=false (homography / Otsu arithmetic follows the dependency's evaluation order).
#include "common.cuh"
#include "chain.cuh"
#include <math.h>
#include <float.h>

#define DEC_THREADS 1024
#define DEC_WARPS (DEC_THREADS / 32)
#define DEC_MAX_S 64                      // canonical image side limit ((markerSize + 2*border) * cellSize)
#define DEC_SMEMC 512                     // candidates per frame handled entirely in shared memory
#define DEC_MAXC APSE_MAX_QUADS           // candidates per frame (beyond DEC_SMEMC the arrays live in global scratch)

// per-frame candidate arrays, carved out of shared memory (n <= DEC_SMEMC) or of a global scratch block
struct DecodeArrays {
    float (*c)[8];                        // candidate corners, sorted by perimeter (descending, stable)
    float *perim;
    uint32_t *key;                        // ordering key (cluster index) / scratch
    uint32_t *close_bits;                 // [cap][cap / 32] too-close predicate, bit (i, j)
    uint32_t *pairs;                      // [cap * 8] too-close pairs (i << 16 | j) in row-major order
    float *cxy;                           // [cap][2] candidate centroids (cheap rejection in the pair matrix)
    short *group_id, *next_in_group, *close_next, *parent, *depth, *sel, *sel_of, *dec_id, *use_c, *group_head, *group_tail;
    uint8_t *dec_valid, *dec_rot, *selected, *was, *valid;
    int cap;
};

__host__ __device__ inline size_t decode_arrays_bytes(int cap)
{
    return (size_t)cap * (8 * 4 + 4 + 4 + (cap / 32) * 4 + 8 * 4 + 2 * 4 + 11 * 2 + 5) + 64;
}

__device__ inline void decode_arrays_carve(DecodeArrays &A, unsigned char *base, int cap)
{
    A.cap = cap;
    unsigned char *p = base;
    A.c = reinterpret_cast<float(*)[8]>(p); p += (size_t)cap * 32;
    A.perim = reinterpret_cast<float *>(p); p += (size_t)cap * 4;
    A.key = reinterpret_cast<uint32_t *>(p); p += (size_t)cap * 4;
    A.close_bits = reinterpret_cast<uint32_t *>(p); p += (size_t)cap * (cap / 32) * 4;
    A.pairs = reinterpret_cast<uint32_t *>(p); p += (size_t)cap * 32;
    A.cxy = reinterpret_cast<float *>(p); p += (size_t)cap * 8;
    short **sp[11] = {&A.group_id, &A.next_in_group, &A.close_next, &A.parent, &A.depth, &A.sel, &A.sel_of, &A.dec_id, &A.use_c,
                      &A.group_head, &A.group_tail};
    for (int i = 0; i < 11; i++) { *sp[i] = reinterpret_cast<short *>(p); p += (size_t)cap * 2; }
    uint8_t **bp[5] = {&A.dec_valid, &A.dec_rot, &A.selected, &A.was, &A.valid};
    for (int i = 0; i < 5; i++) { *bp[i] = p; p += cap; }
}

struct DecodeSmem {
    int n, ns, ngroups;
    int pad;
    // followed by decode_arrays_bytes(DEC_SMEMC) bytes
};

__device__ __forceinline__ float sqf(float v) { return v * v; }

__device__ float perimeter_of(const float *c)
{
    float p = 0.f;
    for (int i = 0; i < 4; i++) {
        int j = (i + 1) % 4;
        p += sqrtf(sqf(c[2 * i] - c[2 * j]) + sqf(c[2 * i + 1] - c[2 * j + 1]));
    }
    return p;
}

__device__ float average_distance(const float *m1, const float *m2)
{
    float best = FLT_MAX;
    for (int fc = 0; fc < 4; fc++) {
        float d = 0;
        for (int c = 0; c < 4; c++) {
            int mc = (c + fc) % 4;
            float dx = m1[2 * mc] - m2[2 * c], dy = m1[2 * mc + 1] - m2[2 * c + 1];
            d += dx * dx + dy * dy;
        }
        d /= 4.f;
        best = fminf(best, d);
    }
    return sqrtf(best);
}

__device__ float average_module_size(const float *c, int ms, int bb)
{
    float a = 0.f;
    for (int i = 0; i < 4; i++) {
        int j = (i + 1) % 4;
        float dx = c[2 * i] - c[2 * j], dy = c[2 * i + 1] - c[2 * j + 1];
        a += sqrtf(dx * dx + dy * dy);
    }
    return a / (4.f * (ms + bb * 2));
}

// pointPolygonTest(poly, pt, false) > 0 for a 4-gon of float points
__device__ bool strictly_inside(const float *poly, float px, float py)
{
    int counter = 0;
    float vx = poly[6], vy = poly[7];
    for (int i = 0; i < 4; i++) {
        float v0x = vx, v0y = vy;
        vx = poly[2 * i]; vy = poly[2 * i + 1];
        if ((v0y <= py && vy <= py) || (v0y > py && vy > py) || (v0x < px && vx < px)) {
            if (py == vy && (px == vx || (py == v0y && ((v0x <= px && px <= vx) || (vx <= px && px <= v0x))))) return false;
            continue;
        }
        double dist = (double)(py - v0y) * (vx - v0x) - (double)(px - v0x) * (vy - v0y);
        if (dist == 0) return false;
        if (vy < v0y) dist = -dist;
        counter += dist > 0;
    }
    return (counter % 2) != 0;
}

// getPerspectiveTransform(corners -> canonical square) followed by the 3x3 inverse used by warpPerspective
__device__ void inverse_homography(const float *src, int S, double *Mi)
{
    double a[64], b[8];
    const float dstc[8] = {0.f, 0.f, (float)S - 1, 0.f, (float)S - 1, (float)S - 1, 0.f, (float)S - 1};
    for (int i = 0; i < 64; i++) a[i] = 0;
    for (int i = 0; i < 4; i++) {
        float sx = src[2 * i], sy = src[2 * i + 1], dx = dstc[2 * i], dy = dstc[2 * i + 1];
        a[i * 8 + 0] = a[(i + 4) * 8 + 3] = sx;
        a[i * 8 + 1] = a[(i + 4) * 8 + 4] = sy;
        a[i * 8 + 2] = a[(i + 4) * 8 + 5] = 1;
        // the dependency forms these products on float32 point members before widening to double
        a[i * 8 + 6] = __fmul_rn(-sx, dx);
        a[i * 8 + 7] = __fmul_rn(-sy, dx);
        a[(i + 4) * 8 + 6] = __fmul_rn(-sx, dy);
        a[(i + 4) * 8 + 7] = __fmul_rn(-sy, dy);
        b[i] = dx;
        b[i + 4] = dy;
    }
    bool ok = true;
    for (int i = 0; i < 8 && ok; i++) {  // LU, partial pivoting
        int k = i;
        for (int j = i + 1; j < 8; j++)
            if (fabs(a[j * 8 + i]) > fabs(a[k * 8 + i])) k = j;
        if (fabs(a[k * 8 + i]) < DBL_EPSILON) { ok = false; break; }
        if (k != i) {
            for (int j = i; j < 8; j++) { double t = a[i * 8 + j]; a[i * 8 + j] = a[k * 8 + j]; a[k * 8 + j] = t; }
            double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        double d = -1 / a[i * 8 + i];
        for (int j = i + 1; j < 8; j++) {
            double alpha = a[j * 8 + i] * d;
            for (k = i + 1; k < 8; k++) a[j * 8 + k] += alpha * a[i * 8 + k];
            b[j] += alpha * b[i];
        }
    }
    if (ok)
        for (int i = 7; i >= 0; i--) {
            double s = b[i];
            for (int k = i + 1; k < 8; k++) s -= a[i * 8 + k] * b[k];
            b[i] = s / a[i * 8 + i];
        }
    else
        for (int i = 0; i < 8; i++) b[i] = 0;
    double m[9] = {b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7], 1.};
    double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    if (d == 0) { for (int i = 0; i < 9; i++) Mi[i] = 0; return; }
    d = 1. / d;
    Mi[0] = (m[4] * m[8] - m[5] * m[7]) * d; Mi[1] = (m[2] * m[7] - m[1] * m[8]) * d; Mi[2] = (m[1] * m[5] - m[2] * m[4]) * d;
    Mi[3] = (m[5] * m[6] - m[3] * m[8]) * d; Mi[4] = (m[0] * m[8] - m[2] * m[6]) * d; Mi[5] = (m[2] * m[3] - m[0] * m[5]) * d;
    Mi[6] = (m[3] * m[7] - m[4] * m[6]) * d; Mi[7] = (m[1] * m[6] - m[0] * m[7]) * d; Mi[8] = (m[0] * m[4] - m[1] * m[3]) * d;
}

// _identifyOneCandidate, two phases.
// (1) decode_sample: ALL threads of the CTA.  Nearest-neighbour warp (rint of the FP64 source coordinate, border 0) into img,
//     histogram, inner-region moments.  This is the latency of the candidate (2304 dependent gathers for a 48 x 48 canonical
//     image), so it is spread over the whole CTA.
// SPARSE: the gray buffer holds exact values only on the tiles flagged in S.eflag; every other sample is computed on demand
// from the source frame (same remap + colour chain as the preprocess kernels, bit-identical by construction)
template <bool SPARSE>
__device__ __forceinline__ void decode_sample(const uint8_t *__restrict__ im, int w, int h, const double *Mi, const DeviceParams &P,
                                              const SparseSrc &SS, int frame, uint8_t *img, int *hist, long long &s1, long long &s2)
{
    const int nb = P.marker_size + 2 * P.border_bits, cs = P.cell_size, S = nb * cs;
    const int c0 = cs / 2, c1 = S - cs / 2;
    s1 = 0; s2 = 0;
    for (int p = threadIdx.x; p < S * S; p += blockDim.x) {
        int y = p / S, x = p - y * S;
        double X0 = Mi[1] * y + Mi[2], Y0 = Mi[4] * y + Mi[5], W0 = Mi[7] * y + Mi[8];
        double W = W0 + Mi[6] * x;
        W = W ? 1. / W : 0;
        double fX = (X0 + Mi[0] * x) * W, fY = (Y0 + Mi[3] * x) * W;
        fX = fmin(fmax(fX, (double)INT32_MIN), (double)INT32_MAX);
        fY = fmin(fmax(fY, (double)INT32_MIN), (double)INT32_MAX);
        long long X = __double2ll_rn(fX), Y = __double2ll_rn(fY);
        int v = 0;
        if (X >= 0 && X < w && Y >= 0 && Y < h) {
            if (!SPARSE || SS.eflag[((size_t)frame * SS.th + ((int)Y >> 2)) * SS.tw + ((int)X >> 2)])
                v = im[(size_t)Y * w + X];
            else
                v = exact_gray_px(SS.bgr + (size_t)frame * w * h * 3, SS.mapx, SS.mapy, SS.tables, w, h, (int)X, (int)Y);
        }
        img[p] = (uint8_t)v;
        atomicAdd(&hist[v], 1);
        if (x >= c0 && x < c1 && y >= c0 && y < c1) { s1 += v; s2 += v * v; }
    }
}

// (2) decode_finish: ONE warp.  Otsu (sequential over the 256 bins in the dependency's order), cell majority, border check,
//     Hamming match against the dictionary.  Returns (valid, id, rot) on every lane.
__device__ void decode_finish(const DeviceParams &P, long long s1, long long s2, const uint8_t *__restrict__ dict, const uint8_t *img,
                              const int *hist, uint8_t *bits, bool &valid, int &id, int &rot, int *thr_out = nullptr)
{
    const int lane = threadIdx.x & 31;
    const int nb = P.marker_size + 2 * P.border_bits, cs = P.cell_size, S = nb * cs;
    const int c0 = cs / 2, c1 = S - cs / 2;
    const int cnt = (c1 - c0) * (c1 - c0);
    double mean = (double)s1 / cnt, var = (double)s2 / cnt - mean * mean;
    double sd = sqrt(var > 0 ? var : 0);
    if (sd < P.min_otsu_stddev) {
        for (int i = lane; i < nb * nb; i += 32) bits[i] = mean > 127 ? 1 : 0;
        if (thr_out && lane == 0) *thr_out = -1;   // flat candidate: no Otsu threshold
    } else {
        int thr = 0;
        if (lane == 0) {  // Otsu, sequential over the 256 bins in the dependency's order
            double mu = 0, scale = 1. / (S * S);
            for (int i = 0; i < 256; i++) mu += i * (double)hist[i];
            mu *= scale;
            double mu1 = 0, q1 = 0, max_sigma = 0;
            for (int i = 0; i < 256; i++) {
                double p_i = hist[i] * scale, q2, mu2, sigma;
                mu1 *= q1;
                q1 += p_i;
                q2 = 1. - q1;
                if (fmin(q1, q2) < FLT_EPSILON || fmax(q1, q2) > 1. - FLT_EPSILON) continue;
                mu1 = (mu1 + i * p_i) / q1;
                mu2 = (mu - q1 * mu1) / q2;
                sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
                if (sigma > max_sigma) { max_sigma = sigma; thr = i; }
            }
        }
        thr = __shfl_sync(0xffffffffu, thr, 0);
        if (thr_out && lane == 0) *thr_out = thr;
        const int margin = P.cell_margin_px, inner = cs - 2 * margin;
        for (int cell = lane; cell < nb * nb; cell += 32) {
            int cy = cell / nb, cx = cell - cy * nb, nz = 0;
            for (int yy = 0; yy < inner; yy++)
                for (int xx = 0; xx < inner; xx++) nz += img[(cy * cs + margin + yy) * S + cx * cs + margin + xx] > thr;
            bits[cell] = nz > (inner * inner) / 2;
        }
    }
    __syncwarp();
    // border errors
    int e = 0;
    for (int cell = lane; cell < nb * nb; cell += 32) {
        int cy = cell / nb, cx = cell - cy * nb;
        bool border = cy < P.border_bits || cy >= nb - P.border_bits || cx < P.border_bits || cx >= nb - P.border_bits;
        e += (border && bits[cell]) ? 1 : 0;
    }
    for (int d = 16; d > 0; d >>= 1) e += __shfl_xor_sync(0xffffffffu, e, d);
    valid = false; id = -1; rot = 0;
    if (e > P.max_border_errors) return;
    // candidate bytes (MSB first, rows of inner bits)
    const int ms = P.marker_size, nbits = ms * ms, nbytes = P.nbytes;
    uint8_t cand[8];
    for (int b = 0; b < nbytes; b++) cand[b] = 0;
    {
        int cur_bit = 0, cur_byte = 0;
        for (int i = 0; i < nbits; i++) {
            int y = i / ms, x = i - y * ms;
            cand[cur_byte] = (uint8_t)(cand[cur_byte] << 1);
            if (bits[(y + P.border_bits) * nb + x + P.border_bits]) cand[cur_byte]++;
            if (++cur_bit == 8) { cur_bit = 0; cur_byte++; }
        }
    }
    // first marker (lowest index) whose best rotation is within the correction budget
    int best_m = 0x7fffffff, best_r = 0;
    for (int m = lane; m < P.n_markers; m += 32) {
        int bd = nbits + 1, br = -1;
        for (int r = 0; r < 4; r++) {
            int hd = 0;
            for (int b = 0; b < nbytes; b++) hd += __popc((unsigned)(dict[(m * 4 + r) * nbytes + b] ^ cand[b]));
            if (hd < bd) { bd = hd; br = r; }
        }
        if (bd <= P.max_correction && m < best_m) { best_m = m; best_r = br; }
    }
    int packed = best_m == 0x7fffffff ? 0x7fffffff : (best_m << 2) | best_r;
    for (int d = 16; d > 0; d >>= 1) packed = min(packed, __shfl_xor_sync(0xffffffffu, packed, d));
    if (packed != 0x7fffffff) { valid = true; id = packed >> 2; rot = packed & 3; }
}

// _identifyOneCandidate for every raw quad of the batch: one CTA per candidate at a time, grid = (candidate slots, frames).
// Results (valid | rot << 1 | id << 8) are indexed like the raw quads; k_decode picks them up after its sort.
#define DECB_THREADS 128
template <bool SPARSE>
__global__ void __launch_bounds__(DECB_THREADS) k_decode_bits(const uint8_t *__restrict__ gray, int w, int h, const float *__restrict__ quads,
                                                              const int32_t *__restrict__ counters, DeviceParams P,
                                                              const uint8_t *__restrict__ dict, int32_t *__restrict__ dec_raw, SparseSrc S)
{
    __shared__ uint8_t s_img[DEC_MAX_S * DEC_MAX_S];
    __shared__ int s_hist[256];
    __shared__ uint8_t s_bits[16 * 16];
    __shared__ double s_Mi[9];
    __shared__ long long s_sum[2][DECB_THREADS / 32];
    const int f = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nq = min(counters[f * APSE_COUNTERS + 2], DEC_MAXC);
    const uint8_t *im = gray + (size_t)f * w * h;
    const int nbc = P.marker_size + 2 * P.border_bits;
    for (int i = blockIdx.x; i < nq; i += gridDim.x) {
        if (threadIdx.x == 0) {
            double Mi[9];
            inverse_homography(quads + ((size_t)f * APSE_MAX_QUADS + i) * 8, nbc * P.cell_size, Mi);
            for (int k = 0; k < 9; k++) s_Mi[k] = Mi[k];
        }
        for (int k = threadIdx.x; k < 256; k += DECB_THREADS) s_hist[k] = 0;
        __syncthreads();
        double Mi[9];
        for (int k = 0; k < 9; k++) Mi[k] = s_Mi[k];
        long long s1, s2;
        decode_sample<SPARSE>(im, w, h, Mi, P, S, f, s_img, s_hist, s1, s2);
        for (int d = 16; d > 0; d >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, d); s2 += __shfl_xor_sync(0xffffffffu, s2, d); }
        if (lane == 0) { s_sum[0][wid] = s1; s_sum[1][wid] = s2; }
        __syncthreads();
        if (wid == 0) {
            s1 = 0; s2 = 0;
            for (int k = 0; k < DECB_THREADS / 32; k++) { s1 += s_sum[0][k]; s2 += s_sum[1][k]; }
            bool v; int id, rot;
            decode_finish(P, s1, s2, dict, s_img, s_hist, s_bits, v, id, rot);
            if (lane == 0) dec_raw[(size_t)f * APSE_MAX_QUADS + i] = (v ? 1 : 0) | (rot << 1) | (id << 8);
        }
        __syncthreads();
    }
}

// test tap (a6.A6 / a6.A7 in isolation): one candidate -> canonical image, Otsu threshold, cell bits, identification
__global__ void __launch_bounds__(DECB_THREADS) k_decode_tap(const uint8_t *__restrict__ gray, int w, int h, const float *__restrict__ corners, int n,
                                                             DeviceParams P, const uint8_t *__restrict__ dict, uint8_t *__restrict__ img_out,
                                                             uint8_t *__restrict__ bits_out, int32_t *__restrict__ res_out)
{
    __shared__ uint8_t s_img[DEC_MAX_S * DEC_MAX_S];
    __shared__ int s_hist[256];
    __shared__ uint8_t s_bits[16 * 16];
    __shared__ double s_Mi[9];
    __shared__ long long s_sum[2][DECB_THREADS / 32];
    __shared__ int s_thr;
    const int i = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (i >= n) return;
    const int nbc = P.marker_size + 2 * P.border_bits, S = nbc * P.cell_size;
    if (threadIdx.x == 0) {
        double Mi[9];
        inverse_homography(corners + 8 * i, S, Mi);
        for (int k = 0; k < 9; k++) s_Mi[k] = Mi[k];
        s_thr = -2;
    }
    for (int k = threadIdx.x; k < 256; k += DECB_THREADS) s_hist[k] = 0;
    __syncthreads();
    double Mi[9];
    for (int k = 0; k < 9; k++) Mi[k] = s_Mi[k];
    long long s1, s2;
    decode_sample<false>(gray, w, h, Mi, P, SparseSrc{}, 0, s_img, s_hist, s1, s2);
    for (int d = 16; d > 0; d >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, d); s2 += __shfl_xor_sync(0xffffffffu, s2, d); }
    if (lane == 0) { s_sum[0][wid] = s1; s_sum[1][wid] = s2; }
    __syncthreads();
    if (wid == 0) {
        s1 = 0; s2 = 0;
        for (int k = 0; k < DECB_THREADS / 32; k++) { s1 += s_sum[0][k]; s2 += s_sum[1][k]; }
        bool v; int id, rot;
        decode_finish(P, s1, s2, dict, s_img, s_hist, s_bits, v, id, rot, &s_thr);
        __syncwarp();
        if (lane == 0) { res_out[4 * i] = v ? 1 : 0; res_out[4 * i + 1] = id; res_out[4 * i + 2] = rot; res_out[4 * i + 3] = s_thr; }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < S * S; k += DECB_THREADS) img_out[(size_t)i * DEC_MAX_S * DEC_MAX_S + k] = s_img[k];
    for (int k = threadIdx.x; k < nbc * nbc; k += DECB_THREADS) bits_out[(size_t)i * 256 + k] = s_bits[k];
}

int apse_decode_tap(apse_ctx *ctx, const uint8_t *gray, int w, int h, const float *corners, int n, const DeviceParams &dp, uint8_t *img_out,
                    uint8_t *bits_out, int32_t *res_out, cudaStream_t st)
{
    int nb = dp.marker_size + 2 * dp.border_bits;
    if (nb * dp.cell_size > DEC_MAX_S || nb > 16 || dp.nbytes > 8) CTX_FAIL(ctx, APSE_ERR_UNSUPPORTED, "decode tap: canonical marker image too large");
    KLAUNCH(ctx, KID_DECODE_BITS, st, k_decode_tap<<<n, DECB_THREADS, 0, st>>>(gray, w, h, corners, n, dp, ctx->dict, img_out, bits_out, res_out));
    return APSE_OK;
}

__global__ void __launch_bounds__(DEC_THREADS) k_decode(const uint8_t *__restrict__ gray, int w, int h, const float *__restrict__ quads,
                                                        const uint32_t *__restrict__ quad_order, int32_t *__restrict__ counters,
                                                        DeviceParams P, const int32_t *__restrict__ dec_raw, int skip_decoded_parents,
                                                        unsigned char *__restrict__ big_scratch, size_t big_stride, apse_detections out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DecodeSmem &S = *reinterpret_cast<DecodeSmem *>(smem_raw);
    __shared__ DecodeArrays A;
    const int f = blockIdx.x, tid = threadIdx.x;
    int32_t *cnt = counters + f * APSE_COUNTERS;
    const int nq = min(cnt[2], DEC_MAXC);
    const float *q = quads + (size_t)f * APSE_MAX_QUADS * 8;
    const uint32_t *qo = quad_order + (size_t)f * APSE_MAX_QUADS;
    if (nq > DEC_SMEMC && !big_scratch) {   // host passes the scratch whenever the path can produce this many quads
        if (tid == 0) { out.n_markers[f] = 0; if (out.n_rejected) out.n_rejected[f] = 0; out.status[f] = APSE_ERR_CAPACITY; }
        return;
    }
    if (tid == 0) {
        if (nq <= DEC_SMEMC) decode_arrays_carve(A, smem_raw + sizeof(DecodeSmem), DEC_SMEMC);
        else decode_arrays_carve(A, big_scratch + (size_t)f * big_stride, DEC_MAXC);
    }
    __syncthreads();
    const int rw = A.cap / 32;   // words per close_bits row

    // ---- deterministic order (candidate key), perimeter, stable sort (descending).  n = nq: the border-distance test
    // comes AFTER the grouping (4.13: a border-touching quad still groups and, as a group main, takes its group with it)
    for (int i = tid; i < nq; i += DEC_THREADS) A.key[i] = qo[i];
    __syncthreads();
    for (int i = tid; i < nq; i += DEC_THREADS) {
        uint32_t k = A.key[i];
        int r = 0;
        for (int j = 0; j < nq; j++) r += A.key[j] < k;
        A.sel[i] = (short)r;  // rank in candidate order
    }
    __syncthreads();
    const int n = nq;
    const int32_t *dr = dec_raw + (size_t)f * APSE_MAX_QUADS;
    for (int i = tid; i < nq; i += DEC_THREADS) {
        int r = A.sel[i];
        for (int k = 0; k < 8; k++) A.c[r][k] = q[8 * i + k];
        A.group_id[r] = (short)(dr[i] & 0xff);      // decode result rides along through the two permutations
        A.next_in_group[r] = (short)(dr[i] >> 8);
    }
    __syncthreads();
    for (int i = tid; i < n; i += DEC_THREADS) A.perim[i] = perimeter_of(A.c[i]);
    __syncthreads();
    // stable descending rank; permute through the key / close_bits scratch (ranks first, then a copy pass)
    for (int i = tid; i < n; i += DEC_THREADS) {
        float p = A.perim[i];
        int r = 0;
        for (int j = 0; j < n; j++) { float pj = A.perim[j]; r += (pj > p) || (pj == p && j < i); }
        A.sel_of[i] = (short)r;
    }
    __syncthreads();
    {
        float *tmp = reinterpret_cast<float *>(A.close_bits);   // >= 9 floats per candidate (cap / 32 >= 16 words per row)
        for (int i = tid; i < n; i += DEC_THREADS) {
            for (int k = 0; k < 8; k++) tmp[9 * i + k] = A.c[i][k];
            tmp[9 * i + 8] = A.perim[i];
        }
        __syncthreads();
        for (int i = tid; i < n; i += DEC_THREADS) {
            int r = A.sel_of[i];
            for (int k = 0; k < 8; k++) A.c[r][k] = tmp[9 * i + k];
            A.perim[r] = tmp[9 * i + 8];
            int flags = A.group_id[i];
            A.dec_valid[r] = (uint8_t)(flags & 1); A.dec_rot[r] = (uint8_t)((flags >> 1) & 3); A.dec_id[r] = A.next_in_group[i];
        }
    }
    __syncthreads();

    // ---- too-close predicate matrix: bit (i,j), i < j, set when avgDist(i,j) < perimeter[j] * rate.
    // avgDist^2 is a mean of squared corner distances, so it is never below the squared centroid distance (Jensen):
    // pairs whose centroids are clearly farther apart than the threshold are skipped without the 4-shift evaluation.
    const int words = (n + 31) / 32;
    for (int i = tid; i < n; i += DEC_THREADS) {
        const float *c = A.c[i];
        A.cxy[2 * i] = 0.25f * (c[0] + c[2] + c[4] + c[6]);
        A.cxy[2 * i + 1] = 0.25f * (c[1] + c[3] + c[5] + c[7]);
    }
    __syncthreads();
    for (int p = tid; p < n * words; p += DEC_THREADS) {
        int i = p / words, wj = p - i * words;
        uint32_t bitsw = 0;
        if (wj * 32 + 31 > i) {
            const float cxi = A.cxy[2 * i], cyi = A.cxy[2 * i + 1];
            for (int b = 0; b < 32; b++) {
                int j = wj * 32 + b;
                if (j > i && j < n) {
                    const float thr = A.perim[j] * P.min_marker_distance_rate;
                    const float dx = A.cxy[2 * j] - cxi, dy = A.cxy[2 * j + 1] - cyi;
                    if (dx * dx + dy * dy > thr * thr * 1.02f + 4.f) continue;   // conservative: float rounding of the centroids
                    float md = average_distance(A.c[i], A.c[j]);
                    if (md < thr) bitsw |= 1u << b;
                }
            }
        }
        A.close_bits[(size_t)i * rw + wj] = bitsw;
    }
    __syncthreads();
    // compact the set bits into a row-major pair list: per-row counts, block scan, scatter
    {
        __shared__ int s_part[DEC_THREADS];
        const int chunk = (n + DEC_THREADS - 1) / DEC_THREADS, r0 = tid * chunk, r1 = min(n, r0 + chunk);
        int local = 0;
        for (int i = r0; i < r1; i++) {
            int c = 0;
            for (int wj = i / 32; wj < words; wj++) c += __popc(A.close_bits[(size_t)i * rw + wj]);
            A.key[i] = (uint32_t)c;
            local += c;
        }
        s_part[tid] = local;
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int t = 0; t < DEC_THREADS; t++) { int v = s_part[t]; s_part[t] = acc; acc += v; }
            S.n = acc;   // total number of too-close pairs
        }
        __syncthreads();
        int off = s_part[tid];
        const int pair_cap = A.cap * 8;
        for (int i = r0; i < r1; i++) {
            if (!A.key[i]) continue;
            for (int wj = i / 32; wj < words; wj++) {
                uint32_t m = A.close_bits[(size_t)i * rw + wj];
                while (m) {
                    int j = wj * 32 + __ffs(m) - 1;
                    m &= m - 1;
                    if (off < pair_cap) A.pairs[off] = ((uint32_t)i << 16) | (uint32_t)j;
                    off++;
                }
            }
        }
    }
    for (int i = tid; i < n; i += DEC_THREADS) {
        A.group_id[i] = -1; A.selected[i] = 1; A.next_in_group[i] = -1; A.close_next[i] = -1;
        A.parent[i] = -1; A.depth[i] = 0; A.was[i] = 0; A.valid[i] = 0; A.use_c[i] = (short)i;
    }
    __syncthreads();

    // ---- sequential grouping walk over the pair list (order-dependent by definition)
    if (tid == 0) {
        int ngroups = 0;
        const int npairs = S.n;
        if (npairs > A.cap * 8) cnt[3] = APSE_ERR_CAPACITY;   // reported through status; never truncated silently
        for (int k = 0; k < min(npairs, A.cap * 8); k++) {
            const uint32_t pr = A.pairs[k];
            const int i = pr >> 16, j = pr & 0xffff;
            A.selected[i] = 0; A.selected[j] = 0;
            if (A.group_id[i] < 0 && A.group_id[j] < 0) { A.group_id[i] = A.group_id[j] = (short)ngroups++; }
            else if (A.group_id[i] > -1 && A.group_id[j] == -1) A.group_id[j] = A.group_id[i];
            else if (A.group_id[j] > -1 && A.group_id[i] == -1) A.group_id[i] = A.group_id[j];
        }
        // members of each group in ascending index order (= largest perimeter first)
        for (int g = 0; g < ngroups; g++) { A.group_head[g] = -1; A.group_tail[g] = -1; }
        for (int i = 0; i < n; i++) {
            int g = A.group_id[i];
            if (g < 0) continue;
            if (A.group_head[g] < 0) A.group_head[g] = (short)i; else A.next_in_group[A.group_tail[g]] = (short)i;
            A.group_tail[g] = (short)i;
        }
        for (int g = 0; g < ngroups; g++) {
            int head = A.group_head[g], cur = head, tail_close = -1;
            A.selected[head] = 1;
            for (int id = A.next_in_group[head]; id >= 0; id = A.next_in_group[id]) {
                float dist = average_distance(A.c[id], A.c[cur]);
                float msz = average_module_size(A.c[id], P.marker_size, P.border_bits);
                if (dist > P.min_group_distance * msz) {
                    cur = id;
                    if (tail_close < 0) A.close_next[head] = (short)id; else A.close_next[tail_close] = (short)id;
                    tail_close = id;
                }
            }
        }
        // NB close_next chains start at the group's head; members never head a chain themselves.
        // Border-distance test on the selected (main) candidates only: a main too near the edge is dropped with its group.
        const float d = (float)P.min_distance_to_border;
        int ns = 0;
        for (int i = 0; i < n; i++) {
            if (!A.selected[i]) continue;
            const float *c = A.c[i];
            bool near = false;
            for (int j = 0; j < 4; j++) near |= c[2 * j] < d || c[2 * j + 1] < d || c[2 * j] > w - 1 - d || c[2 * j + 1] > h - 1 - d;
            if (near) continue;
            A.sel[ns] = (short)i; A.sel_of[i] = (short)ns; ns++;
        }
        S.ns = ns;
        S.ngroups = ngroups;
    }
    __syncthreads();
    const int ns = S.ns;

    // ---- nesting hierarchy among the selected candidates: parent = nearest smaller index that contains all 4 corners
    for (int v = tid; v < ns; v += DEC_THREADS) {
        const float *a = A.c[A.sel[v]];
        int par = -1;
        for (int j = v - 1; j >= 0; j--) {
            const float *b = A.c[A.sel[j]];
            if (strictly_inside(b, a[0], a[1]) && strictly_inside(b, a[2], a[3]) && strictly_inside(b, a[4], a[5]) &&
                strictly_inside(b, a[6], a[7])) { par = j; break; }
        }
        A.parent[v] = (short)par;
    }
    __syncthreads();
    if (tid == 0) {
        int max_depth = 0;
        for (int v = ns - 1; v >= 0; v--) {
            int p = A.parent[v];
            if (p >= 0 && A.depth[v] + 1 > A.depth[p]) A.depth[p] = (short)(A.depth[v] + 1);
        }
        for (int v = 0; v < ns; v++) max_depth = max(max_depth, (int)A.depth[v]);
        int counter = 0;
        for (int dep = 0; dep <= max_depth && counter < ns; dep++) {
            for (int v = 0; v < ns; v++) {
                if (A.depth[v] != dep) continue;
                if (skip_decoded_parents && A.was[v]) continue;
                A.was[v] = 1;
                int head = A.sel[v];
                int use = -1;
                if (A.dec_valid[head]) use = head;
                else
                    for (int c = A.close_next[head]; c >= 0; c = A.close_next[c])
                        if (A.dec_valid[c]) { use = c; break; }
                if (use >= 0) { A.valid[v] = 1; A.use_c[v] = (short)use; }
            }
            for (int v = 0; v < ns; v++) {
                if (A.depth[v] != dep) continue;
                if (A.valid[v]) {
                    int p = A.parent[v];
                    while (p != -1) {
                        if (!A.was[p]) { A.was[p] = 1; counter++; }
                        p = A.parent[p];
                    }
                }
                counter++;
            }
        }
        // ---- output
        int na = 0, nr = 0;
        const int cap = out.max_markers;
        float *oc = out.corners + (size_t)f * cap * 8;
        int32_t *oi = out.ids + (size_t)f * cap;
        float *orj = out.rejected ? out.rejected + (size_t)f * cap * 8 : nullptr;
        int status = cnt[3];
        for (int v = 0; v < ns; v++) {
            if (A.valid[v]) {
                int u = A.use_c[v];
                const float *c = A.c[u];
                if (na < cap) {
                    int r = A.dec_rot[u];
                    for (int k = 0; k < 4; k++) {
                        int s = (k + 4 - r) % 4;
                        oc[8 * na + 2 * k] = c[2 * s];
                        oc[8 * na + 2 * k + 1] = c[2 * s + 1];
                    }
                    oi[na] = A.dec_id[u];
                } else status = APSE_ERR_CAPACITY;
                na++;
            } else {
                const float *c = A.c[A.sel[v]];
                if (orj) {
                    if (nr < cap) for (int k = 0; k < 8; k++) orj[8 * nr + k] = c[k];
                    else status = APSE_ERR_CAPACITY;
                }
                nr++;
            }
        }
        out.n_markers[f] = min(na, cap);
        if (out.n_rejected) out.n_rejected[f] = min(nr, cap);
        out.status[f] = status;
    }
}

int apse_decode_alloc(apse_ctx *ctx)
{
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_decode, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(sizeof(DecodeSmem) + decode_arrays_bytes(DEC_SMEMC))));
    return APSE_OK;
}
void apse_decode_free(apse_ctx *ctx) { cudaFree(ctx->decode_scratch); ctx->decode_scratch = nullptr; }

// global candidate arrays for frames with more than DEC_SMEMC quads (classic path); allocated on first use
int apse_decode_big_scratch(apse_ctx *ctx)
{
    if (!ctx->decode_scratch) CUDA_TRY(ctx, cudaMalloc(&ctx->decode_scratch, decode_arrays_bytes(DEC_MAXC) * ctx->max_batch));
    return APSE_OK;
}

int apse_decode_candidates(apse_ctx *ctx, const uint8_t *gray, int w, int h, int batch, const DeviceParams &dp,
                           apse_detections *out, cudaStream_t st)
{
    if (!out || !out->corners || !out->ids || !out->n_markers || !out->status || out->max_markers <= 0)
        CTX_FAIL(ctx, APSE_ERR_INVALID_ARG, "detect: incomplete apse_detections");
    if (out->rejected && !out->n_rejected) CTX_FAIL(ctx, APSE_ERR_INVALID_ARG, "detect: rejected without n_rejected");
    int nb = dp.marker_size + 2 * dp.border_bits;
    if (nb * dp.cell_size > DEC_MAX_S || nb > 16 || dp.nbytes > 8)
        CTX_FAIL(ctx, APSE_ERR_UNSUPPORTED, "detect: canonical marker image %d px exceeds the %d px limit", nb * dp.cell_size, DEC_MAX_S);
    // a quad that encloses an already decoded marker IS identified (cv2 4.13 on nested-marker frames, tests/test_gpu_detect.py);
    // development switch APSE_SKIP_DECODED_PARENTS=1 restores the behaviour round 1 assumed
    static const bool skip_env = getenv("APSE_SKIP_DECODED_PARENTS") != nullptr && getenv("APSE_SKIP_DECODED_PARENTS")[0] == '1';
    int skip = skip_env ? 1 : 0;
    // hierarchy scratch of the quad fit is free at this point: [batch][APSE_MAX_QUADS] decode results
    int32_t *dec_raw = reinterpret_cast<int32_t *>(ctx->errs);
    const int cand_blocks = ctx->params.cornerRefinementMethod == 3 ? 64 : 512;   // CTAs per frame; a CTA without a candidate exits at once
    if (ctx->sparse_active)
        KLAUNCH(ctx, KID_DECODE_BITS, st, k_decode_bits<true><<<dim3(cand_blocks, batch), DECB_THREADS, 0, st>>>(gray, w, h, ctx->quads, ctx->counters, dp, ctx->dict, dec_raw, ctx->sparse_src));
    else
        KLAUNCH(ctx, KID_DECODE_BITS, st, k_decode_bits<false><<<dim3(cand_blocks, batch), DECB_THREADS, 0, st>>>(gray, w, h, ctx->quads, ctx->counters, dp, ctx->dict, dec_raw, SparseSrc{}));
    KLAUNCH(ctx, KID_DECODE, st, k_decode<<<batch, DEC_THREADS, sizeof(DecodeSmem) + decode_arrays_bytes(DEC_SMEMC), st>>>(
                gray, w, h, ctx->quads, ctx->quad_order, ctx->counters, dp, dec_raw, skip,
                (unsigned char *)ctx->decode_scratch, decode_arrays_bytes(DEC_MAXC), *out));
    return APSE_OK;
}
