/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the candidate post-filter, bit extraction and
 * dictionary identification that cv2.aruco (4.13 semantics) applies to quad candidates inside the
 * aruco.detectMarkers call of aruco_detect.py:267 (dictionary from aruco_detect.py:263).
 *
 * Restated from SURVEY.md section 8 rows a6.A5 - a6.A7 / Appendix A.4, A.4b, A.9 (OpenCV is an un-vendored
 * dependency, reference README.md:42); pinned against the cv2 4.13.0 binary by tests/test_oracle_detect.py.
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int marker_size;          /* dictionary.markerSize (4)                     */
    int border_bits;          /* markerBorderBits                              */
    int cell_size;            /* perspectiveRemovePixelPerCell                 */
    double cell_margin_rate;  /* perspectiveRemoveIgnoredMarginPerCell         */
    double min_otsu_stddev;   /* minOtsuStdDev                                 */
    double max_border_err_rate; /* maxErroneousBitsInBorderRate                */
    double error_correction_rate;
    int max_correction_bits;  /* dictionary.maxCorrectionBits                  */
    int min_distance_to_border;
    double min_marker_distance_rate;
    float min_group_distance;
    int detect_inverted;
    int skip_decoded_parents; /* 1: a quad enclosing an already decoded marker is not identified.  cv2 4.13 DOES identify such
                               * quads (24 nested-marker frames, tests/test_oracle_detect.py): 0 is the pinned behaviour */
} orc_dec_params;

/* ------------------------------------------------------------------------------------------------------ */
/* 8x8 linear solve, LU with partial pivoting (getPerspectiveTransform) */
static int solve_lu(double *A, double *b, int n)
{
    for (int i = 0; i < n; i++) {
        int k = i;
        for (int j = i + 1; j < n; j++)
            if (fabs(A[j * n + i]) > fabs(A[k * n + i])) k = j;
        if (fabs(A[k * n + i]) < DBL_EPSILON) return 0;
        if (k != i) {
            for (int j = i; j < n; j++) { double t = A[i * n + j]; A[i * n + j] = A[k * n + j]; A[k * n + j] = t; }
            double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        double d = -1 / A[i * n + i];
        for (int j = i + 1; j < n; j++) {
            double alpha = A[j * n + i] * d;
            for (k = i + 1; k < n; k++) A[j * n + k] += alpha * A[i * n + k];
            b[j] += alpha * b[i];
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = b[i];
        for (int k = i + 1; k < n; k++) s -= A[i * n + k] * b[k];
        b[i] = s / A[i * n + i];
    }
    return 1;
}

void orc_perspective_transform(const float *src, const float *dst, double *M)
{
    double a[64], b[8];
    memset(a, 0, sizeof a);
    for (int i = 0; i < 4; i++) {
        float sx = src[2 * i], sy = src[2 * i + 1], dx = dst[2 * i], dy = dst[2 * i + 1];
        a[i * 8 + 0] = a[(i + 4) * 8 + 3] = sx;
        a[i * 8 + 1] = a[(i + 4) * 8 + 4] = sy;
        a[i * 8 + 2] = a[(i + 4) * 8 + 5] = 1;
        /* the dependency forms these products on float32 point members before widening to double */
        a[i * 8 + 6] = (float)(-sx * dx);
        a[i * 8 + 7] = (float)(-sy * dx);
        a[(i + 4) * 8 + 6] = (float)(-sx * dy);
        a[(i + 4) * 8 + 7] = (float)(-sy * dy);
        b[i] = dx;
        b[i + 4] = dy;
    }
    if (!solve_lu(a, b, 8)) memset(b, 0, sizeof b);
    memcpy(M, b, sizeof b);
    M[8] = 1.;
}

static void invert3(const double *m, double *o)
{
    double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    if (d == 0) { memset(o, 0, 9 * sizeof(double)); return; }
    d = 1. / d;
    o[0] = (m[4] * m[8] - m[5] * m[7]) * d;
    o[1] = (m[2] * m[7] - m[1] * m[8]) * d;
    o[2] = (m[1] * m[5] - m[2] * m[4]) * d;
    o[3] = (m[5] * m[6] - m[3] * m[8]) * d;
    o[4] = (m[0] * m[8] - m[2] * m[6]) * d;
    o[5] = (m[2] * m[3] - m[0] * m[5]) * d;
    o[6] = (m[3] * m[7] - m[4] * m[6]) * d;
    o[7] = (m[1] * m[6] - m[0] * m[7]) * d;
    o[8] = (m[0] * m[4] - m[1] * m[3]) * d;
}

/* warpPerspective(INTER_NEAREST, BORDER_CONSTANT 0) of the quad into an S x S canonical image */
void orc_warp_nearest(const uint8_t *im, int w, int h, const float *corners, int S, uint8_t *out)
{
    float dst[8] = {0, 0, (float)S - 1, 0, (float)S - 1, (float)S - 1, 0, (float)S - 1};
    double M[9], Mi[9];
    orc_perspective_transform(corners, dst, M);
    invert3(M, Mi);
    for (int y = 0; y < S; y++) {
        double X0 = Mi[1] * y + Mi[2], Y0 = Mi[4] * y + Mi[5], W0 = Mi[7] * y + Mi[8];
        for (int x = 0; x < S; x++) {
            double W = W0 + Mi[6] * x;
            W = W ? 1. / W : 0;
            double fX = (X0 + Mi[0] * x) * W, fY = (Y0 + Mi[3] * x) * W;
            if (fX < (double)INT32_MIN) fX = (double)INT32_MIN;
            if (fX > (double)INT32_MAX) fX = (double)INT32_MAX;
            if (fY < (double)INT32_MIN) fY = (double)INT32_MIN;
            if (fY > (double)INT32_MAX) fY = (double)INT32_MAX;
            long X = lrint(fX), Y = lrint(fY);
            out[y * S + x] = (X >= 0 && X < w && Y >= 0 && Y < h) ? im[(size_t)Y * w + X] : 0;
        }
    }
}

int orc_otsu(const uint8_t *px, int n)
{
    int hist[256] = {0};
    for (int i = 0; i < n; i++) hist[px[i]]++;
    double mu = 0, scale = 1. / n;
    for (int i = 0; i < 256; i++) mu += i * (double)hist[i];
    mu *= scale;
    double mu1 = 0, q1 = 0, max_sigma = 0;
    int max_val = 0;
    for (int i = 0; i < 256; i++) {
        double p_i = hist[i] * scale, q2, mu2, sigma;
        mu1 *= q1;
        q1 += p_i;
        q2 = 1. - q1;
        if (fmin(q1, q2) < FLT_EPSILON || fmax(q1, q2) > 1. - FLT_EPSILON) continue;
        mu1 = (mu1 + i * p_i) / q1;
        mu2 = (mu - q1 * mu1) / q2;
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
        if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
    }
    return max_val;
}

/* _extractBits: bits[(ms+2b)^2], 1 = white */
void orc_extract_bits(const uint8_t *im, int w, int h, const float *corners, const orc_dec_params *P, uint8_t *bits,
                      uint8_t *warp_out /* nullable, S*S */, int *otsu_out /* nullable */)
{
    int n = P->marker_size + 2 * P->border_bits;
    int cs = P->cell_size, S = n * cs;
    int margin = (int)(P->cell_margin_rate * cs);
    uint8_t *img = (uint8_t *)malloc((size_t)S * S);
    orc_warp_nearest(im, w, h, corners, S, img);
    if (warp_out) memcpy(warp_out, img, (size_t)S * S);
    if (otsu_out) *otsu_out = -1;
    /* mean / stddev of the inner region (cell_size/2 cropped on each side) */
    int c0 = cs / 2, c1 = S - cs / 2;
    double s = 0, s2 = 0;
    int cnt = 0;
    for (int y = c0; y < c1; y++)
        for (int x = c0; x < c1; x++) {
            double v = img[y * S + x];
            s += v; s2 += v * v; cnt++;
        }
    double mean = s / cnt, var = s2 / cnt - mean * mean;
    double sd = sqrt(var > 0 ? var : 0);
    if (sd < P->min_otsu_stddev) {
        memset(bits, mean > 127 ? 1 : 0, (size_t)n * n);
        free(img);
        return;
    }
    int thr = orc_otsu(img, S * S);
    if (otsu_out) *otsu_out = thr;
    int inner = cs - 2 * margin;
    for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) {
            int nz = 0;
            for (int yy = 0; yy < inner; yy++)
                for (int xx = 0; xx < inner; xx++)
                    nz += img[(y * cs + margin + yy) * S + x * cs + margin + xx] > thr;
            bits[y * n + x] = nz > (inner * inner) / 2;
        }
    free(img);
}

static int border_errors(const uint8_t *bits, int ms, int bb)
{
    int n = ms + 2 * bb, e = 0;
    for (int y = 0; y < n; y++)
        for (int k = 0; k < bb; k++) {
            if (bits[y * n + k]) e++;
            if (bits[y * n + n - 1 - k]) e++;
        }
    for (int x = bb; x < n - bb; x++)
        for (int k = 0; k < bb; k++) {
            if (bits[k * n + x]) e++;
            if (bits[(n - 1 - k) * n + x]) e++;
        }
    return e;
}

/* Dictionary::identify. bytes_list: [n_markers][4 rotations][nbytes] (rotation-major rows) */
int orc_identify(const uint8_t *inner_bits, int ms, const uint8_t *bytes_list, int n_markers, int max_corr_bits,
                 double rate, int *idx, int *rotation)
{
    int nbits = ms * ms, nbytes = (nbits + 7) / 8;
    uint8_t cand[16] = {0};
    {
        int cur_bit = 0, cur_byte = 0;
        for (int i = 0; i < nbits; i++) {
            cand[cur_byte] = (uint8_t)(cand[cur_byte] << 1);
            if (inner_bits[i]) cand[cur_byte]++;
            cur_bit++;
            if (cur_bit == 8) { cur_bit = 0; cur_byte++; }
        }
    }
    int max_corr = (int)((double)max_corr_bits * rate);
    *idx = -1;
    for (int m = 0; m < n_markers; m++) {
        int best = nbits + 1, rot = -1;
        for (int r = 0; r < 4; r++) {
            int hd = 0;
            for (int b = 0; b < nbytes; b++) hd += __builtin_popcount(bytes_list[(m * 4 + r) * nbytes + b] ^ cand[b]);
            if (hd < best) { best = hd; rot = r; }
        }
        if (best <= max_corr) { *idx = m; *rotation = rot; break; }
    }
    return *idx != -1;
}

/* _identifyOneCandidate */
static int identify_one(const uint8_t *im, int w, int h, const float *corners, const orc_dec_params *P,
                        const uint8_t *bytes_list, int n_markers, int *idx, int *rot)
{
    int n = P->marker_size + 2 * P->border_bits;
    uint8_t bits[32 * 32];
    orc_extract_bits(im, w, h, corners, P, bits, NULL, NULL);
    int max_border = (int)(P->marker_size * P->marker_size * P->max_border_err_rate);
    int berr = border_errors(bits, P->marker_size, P->border_bits);
    if (P->detect_inverted) {   /* detectInvertedMarker: a white marker is read through the inverted bits when its border fits better */
        uint8_t inv[32 * 32];
        for (int i = 0; i < n * n; i++) inv[i] = (uint8_t)!bits[i];
        int ierr = border_errors(inv, P->marker_size, P->border_bits);
        if (ierr < berr) { berr = ierr; memcpy(bits, inv, (size_t)n * n); }
    }
    if (berr > max_border) return 0;
    uint8_t inner[32 * 32];
    for (int y = 0; y < P->marker_size; y++)
        for (int x = 0; x < P->marker_size; x++)
            inner[y * P->marker_size + x] = bits[(y + P->border_bits) * n + x + P->border_bits];
    return orc_identify(inner, P->marker_size, bytes_list, n_markers, P->max_correction_bits,
                        P->error_correction_rate, idx, rot);
}

/* ------------------------------------------------------------------------------------------------------ */
/* candidate post-filter (4.13): border distance, stable perimeter sort, too-close grouping, hierarchy */
static float sqf(float v) { return v * v; }

static float perimeter_of(const float *c)
{
    float p = 0.f;
    for (int i = 0; i < 4; i++) {
        int j = (i + 1) % 4;
        p += sqrtf(sqf(c[2 * i] - c[2 * j]) + sqf(c[2 * i + 1] - c[2 * j + 1]));
    }
    return p;
}

static float average_distance(const float *m1, const float *m2)
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
        if (d < best) best = d;
    }
    return sqrtf(best);
}

static float average_module_size(const float *c, int ms, int bb)
{
    float a = 0.f;
    for (int i = 0; i < 4; i++) {
        int j = (i + 1) % 4;
        float dx = c[2 * i] - c[2 * j], dy = c[2 * i + 1] - c[2 * j + 1];
        a += sqrtf(dx * dx + dy * dy);
    }
    return a / (4.f * (ms + bb * 2));
}

/* pointPolygonTest(poly, pt, false) > 0 */
static int strictly_inside(const float *poly, float px, float py)
{
    int counter = 0;
    float vx = poly[6], vy = poly[7];
    for (int i = 0; i < 4; i++) {
        float v0x = vx, v0y = vy;
        vx = poly[2 * i]; vy = poly[2 * i + 1];
        if ((v0y <= py && vy <= py) || (v0y > py && vy > py) || (v0x < px && vx < px)) {
            if (py == vy && (px == vx || (py == v0y && ((v0x <= px && px <= vx) || (vx <= px && px <= v0x))))) return 0;
            continue;
        }
        double dist = (double)(py - v0y) * (vx - v0x) - (double)(px - v0x) * (vy - v0y);
        if (dist == 0) return 0;
        if (vy < v0y) dist = -dist;
        counter += dist > 0;
    }
    return counter % 2 != 0;
}

typedef struct {
    float c[8];
    float perimeter;
    int order;
    int parent, depth;
    int n_close, close_cap;
    int *close; /* indices into the sorted candidate array */
} cand_t;

static int cmp_cand(const void *a, const void *b)
{
    const cand_t *p = (const cand_t *)a, *q = (const cand_t *)b;
    if (p->perimeter != q->perimeter) return p->perimeter > q->perimeter ? -1 : 1;
    return p->order - q->order; /* stable */
}
static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }
static int cmp_int_desc(const void *a, const void *b) { return *(const int *)b - *(const int *)a; }

/* Full candidate -> marker stage.  quads: [nq][8] in detection order.
 * Outputs: corners [max][8] (rotated to the marker's own top-left), ids [max], rejected [max][8].
 * Returns number of accepted markers; *n_rejected receives the rejected count. */
int orc_identify_candidates(const uint8_t *im, int w, int h, const float *quads, int nq, const orc_dec_params *P,
                            const uint8_t *bytes_list, int n_markers, float *corners_out, int32_t *ids_out,
                            float *rejected_out, int max_out, int *n_rejected)
{
    cand_t *cand = (cand_t *)calloc(nq > 0 ? nq : 1, sizeof(cand_t));
    int n = 0;
    float d = (float)P->min_distance_to_border;
    for (int i = 0; i < nq; i++) {
        const float *q = quads + 8 * i;
        memcpy(cand[n].c, q, sizeof(float) * 8);
        cand[n].perimeter = perimeter_of(q);
        cand[n].order = n;
        cand[n].parent = -1;
        n++;
    }
    qsort(cand, n, sizeof(cand_t), cmp_cand);
    int *group_id = (int *)malloc(sizeof(int) * (n + 1));
    uint8_t *selected = (uint8_t *)malloc(n + 1);
    int **groups = (int **)calloc(n + 1, sizeof(int *));
    int *gsize = (int *)calloc(n + 1, sizeof(int)), *gcap = (int *)calloc(n + 1, sizeof(int));
    int ngroups = 0;
    for (int i = 0; i < n; i++) { group_id[i] = -1; selected[i] = 1; }
#define GPUSH(g, v) do { if (gsize[g] == gcap[g]) { gcap[g] = gcap[g] ? 2 * gcap[g] : 4; groups[g] = (int *)realloc(groups[g], sizeof(int) * gcap[g]); } groups[g][gsize[g]++] = (v); } while (0)
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) {
            float md = average_distance(cand[i].c, cand[j].c);
            if (md < cand[j].perimeter * (float)P->min_marker_distance_rate) {
                selected[i] = selected[j] = 0;
                if (group_id[i] < 0 && group_id[j] < 0) {
                    group_id[i] = group_id[j] = ngroups;
                    GPUSH(ngroups, i); GPUSH(ngroups, j);
                    ngroups++;
                } else if (group_id[i] > -1 && group_id[j] == -1) {
                    group_id[j] = group_id[i];
                    GPUSH(group_id[i], j);
                } else if (group_id[j] > -1 && group_id[i] == -1) {
                    group_id[i] = group_id[j];
                    GPUSH(group_id[j], i);
                }
            }
        }
    for (int g = 0; g < ngroups; g++) {
        qsort(groups[g], gsize[g], sizeof(int), P->detect_inverted ? cmp_int_desc : cmp_int);
        int cur = groups[g][0], head = groups[g][0];
        selected[cur] = 1;
        for (int k = 1; k < gsize[g]; k++) {
            int id = groups[g][k];
            float dist = average_distance(cand[id].c, cand[cur].c);
            float msz = average_module_size(cand[id].c, P->marker_size, P->border_bits);
            if (dist > P->min_group_distance * msz) {
                cur = id;
                if (cand[head].n_close == cand[head].close_cap) {
                    cand[head].close_cap = cand[head].close_cap ? 2 * cand[head].close_cap : 4;
                    cand[head].close = (int *)realloc(cand[head].close, sizeof(int) * cand[head].close_cap);
                }
                cand[head].close[cand[head].n_close++] = id;
            }
        }
    }
    /* compact the selected candidates (sorted order kept) */
    int *sel = (int *)malloc(sizeof(int) * (n + 1));
    int ns = 0;
    /* 4.13: the border-distance test is applied AFTER the grouping, to the group mains only: a main closer than
     * minDistanceToBorder to the image edge disappears together with its close contours (neither accepted nor
     * rejected); border-touching quads therefore still take part in the grouping above.  Pinned by
     * tests/test_oracle_classic.py::test_border_quad_swallows_group. */
    for (int i = 0; i < n; i++) {
        if (!selected[i]) continue;
        const float *q = cand[i].c;
        int near = 0;
        for (int j = 0; j < 4; j++)
            if (q[2 * j] < d || q[2 * j + 1] < d || q[2 * j] > w - 1 - d || q[2 * j + 1] > h - 1 - d) near = 1;
        if (!near) sel[ns++] = i;
    }
    int *parent = (int *)malloc(sizeof(int) * (ns + 1)), *depth = (int *)calloc(ns + 1, sizeof(int));
    for (int i = 0; i < ns; i++) parent[i] = -1;
    for (int i = ns - 1; i >= 0; i--)
        for (int j = i - 1; j >= 0; j--) {
            const float *a = cand[sel[i]].c, *b = cand[sel[j]].c;
            if (strictly_inside(b, a[0], a[1]) && strictly_inside(b, a[2], a[3]) && strictly_inside(b, a[4], a[5]) &&
                strictly_inside(b, a[6], a[7])) {
                parent[i] = j;
                if (depth[i] + 1 > depth[j]) depth[j] = depth[i] + 1;
                break;
            }
        }
    int *ids = (int *)malloc(sizeof(int) * (ns + 1)), *rots = (int *)calloc(ns + 1, sizeof(int));
    uint8_t *valid = (uint8_t *)calloc(ns + 1, 1), *was = (uint8_t *)calloc(ns + 1, 1);
    float *final_c = (float *)malloc(sizeof(float) * 8 * (ns + 1));
    int max_depth = 0;
    for (int i = 0; i < ns; i++) {
        ids[i] = -1;
        memcpy(final_c + 8 * i, cand[sel[i]].c, sizeof(float) * 8);
        if (depth[i] > max_depth) max_depth = depth[i];
    }
    int counter = 0;
    for (int dep = 0; dep <= max_depth && counter < ns; dep++) {
        for (int v = 0; v < ns; v++) {
            if (depth[v] != dep) continue;
            if (P->skip_decoded_parents && was[v]) continue;
            was[v] = 1;
            const cand_t *cd = &cand[sel[v]];
            valid[v] = (uint8_t)identify_one(im, w, h, cd->c, P, bytes_list, n_markers, &ids[v], &rots[v]);
            if (!valid[v]) {
                for (int k = 0; k < cd->n_close; k++) {
                    const float *cc = cand[cd->close[k]].c;
                    valid[v] = (uint8_t)identify_one(im, w, h, cc, P, bytes_list, n_markers, &ids[v], &rots[v]);
                    if (valid[v]) { memcpy(final_c + 8 * v, cc, sizeof(float) * 8); break; }
                }
            }
        }
        for (int v = 0; v < ns; v++) {
            if (depth[v] != dep) continue;
            if (valid[v]) {
                int p = parent[v];
                while (p != -1) {
                    if (!was[p]) { was[p] = 1; counter++; }
                    p = parent[p];
                }
            }
            counter++;
        }
    }
    int na = 0, nr = 0;
    for (int v = 0; v < ns; v++) {
        const float *c = final_c + 8 * v;
        if (valid[v]) {
            if (na < max_out) {
                int r = rots[v]; /* std::rotate(begin, begin + 4 - r, end) */
                for (int k = 0; k < 4; k++) {
                    int s = (k + 4 - r) % 4;
                    corners_out[8 * na + 2 * k] = c[2 * s];
                    corners_out[8 * na + 2 * k + 1] = c[2 * s + 1];
                }
                ids_out[na] = ids[v];
            }
            na++;
        } else {
            if (nr < max_out) memcpy(rejected_out + 8 * nr, c, sizeof(float) * 8);
            nr++;
        }
    }
    *n_rejected = nr;
    for (int i = 0; i < n; i++) free(cand[i].close);
    for (int g = 0; g < ngroups; g++) free(groups[g]);
    free(cand); free(group_id); free(selected); free(groups); free(gsize); free(gcap); free(sel); free(parent);
    free(depth); free(ids); free(rots); free(valid); free(was); free(final_c);
    return na;
}
