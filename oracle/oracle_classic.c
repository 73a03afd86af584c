/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the CLASSIC candidate stage that cv2.aruco (4.13
 * semantics) runs inside aruco.detectMarkers (aruco_detect.py:267) when cornerRefinementMethod is NONE / SUBPIX:
 *   adaptiveThreshold (MEAN_C, BINARY_INV) per window  ->  findContours(RETR_LIST, CHAIN_APPROX_NONE)  ->
 *   approxPolyDP / convexity / size / border filters  ->  clockwise reorder  ->  (shared) grouping + decoding
 *   -> cornerSubPix on the accepted markers.
 * north_star stages (2)-(4), BASELINE.json config 5; SURVEY.md section 8 rows a6.C1-a6.C4, Appendix A.5-A.7.
 * OpenCV is an un-vendored dependency of the reference (README.md:42): this file restates its published
 * algorithms (box-mean threshold, Suzuki-Abe border following, Douglas-Peucker, Foerstner-style sub-pixel
 * refinement) and is pinned against the cv2 4.13.0 binary by tests/test_oracle_classic.py.
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int win_min, win_max, win_step;      /* adaptiveThreshWinSize{Min,Max,Step} */
    double constant;                     /* adaptiveThreshConstant              */
    double min_perimeter_rate, max_perimeter_rate;
    double approx_accuracy_rate, min_corner_distance_rate;
    int min_distance_to_border;
} orc_classic_params;

/* ------------------------------------------------------------------------------------------------------ */
/* a6.C1: box mean (BORDER_REPLICATE) rounded half-to-even, out = 255 iff src - mean <= -floor(C) */
static int rint_div(long long s, double inv)
{
    return (int)nearbyint((double)s * inv);
}

void orc_adaptive_threshold(const uint8_t *g, int w, int h, int win, double C, uint8_t *out)
{
    if (win % 2 == 0) win++;
    int r = win / 2;
    double inv = 1.0 / ((double)win * win);
    int idelta = (int)floor(C);
    /* horizontal sums per row, then vertical sliding sum */
    int32_t *hs = (int32_t *)malloc((size_t)w * h * sizeof(int32_t));
    for (int y = 0; y < h; y++) {
        const uint8_t *row = g + (size_t)y * w;
        int32_t *o = hs + (size_t)y * w;
        int s = 0;
        for (int k = -r; k <= r; k++) { int x = k < 0 ? 0 : k >= w ? w - 1 : k; s += row[x]; }
        o[0] = s;
        for (int x = 1; x < w; x++) {
            int xa = x + r >= w ? w - 1 : x + r, xs = x - r - 1 < 0 ? 0 : x - r - 1;
            s += row[xa] - row[xs];
            o[x] = s;
        }
    }
    int32_t *col = (int32_t *)calloc(w, sizeof(int32_t));
    for (int k = -r; k <= r; k++) {
        int y = k < 0 ? 0 : k >= h ? h - 1 : k;
        for (int x = 0; x < w; x++) col[x] += hs[(size_t)y * w + x];
    }
    for (int y = 0; y < h; y++) {
        if (y > 0) {
            int ya = y + r >= h ? h - 1 : y + r, ys = y - r - 1 < 0 ? 0 : y - r - 1;
            for (int x = 0; x < w; x++) col[x] += hs[(size_t)ya * w + x] - hs[(size_t)ys * w + x];
        }
        for (int x = 0; x < w; x++) {
            int mean = rint_div(col[x], inv);
            if (mean > 255) mean = 255;
            out[(size_t)y * w + x] = ((int)g[(size_t)y * w + x] - mean <= -idelta) ? 255 : 0;
        }
    }
    free(col);
    free(hs);
}

/* ------------------------------------------------------------------------------------------------------ */
/* a6.C2: Suzuki-Abe border following on a zero-padded label image (Appendix A.5).
 * Directions 0..7 = E, NE, N, NW, W, SW, S, SE (y grows downwards). */
static const int DX[8] = {1, 1, 0, -1, -1, -1, 0, 1};
static const int DY[8] = {0, -1, -1, -1, 0, 1, 1, 1};

typedef struct {
    int32_t *pts;      /* x,y pairs of all borders, in discovery order */
    size_t n_pts, cap_pts;
    size_t *start;     /* first point of border i */
    int n, cap;
} border_list;

static void bl_push_pt(border_list *L, int x, int y)
{
    if (L->n_pts == L->cap_pts) {
        L->cap_pts = L->cap_pts ? L->cap_pts * 2 : 1 << 16;
        L->pts = (int32_t *)realloc(L->pts, L->cap_pts * 2 * sizeof(int32_t));
    }
    L->pts[2 * L->n_pts] = x;
    L->pts[2 * L->n_pts + 1] = y;
    L->n_pts++;
}

static void bl_begin(border_list *L)
{
    if (L->n == L->cap) {
        L->cap = L->cap ? L->cap * 2 : 1 << 12;
        L->start = (size_t *)realloc(L->start, ((size_t)L->cap + 1) * sizeof(size_t));
    }
    L->start[L->n++] = L->n_pts;
}

/* traces one border starting at (x0,y0) of the padded image f (pitch fw); outer: first search starts from W */
static void trace_border(int32_t *f, int fw, int x0, int y0, int hole, int nbd, border_list *L)
{
    int s = hole ? 0 : 4, s_end = s;
    int found = 0;
    do {
        s = (s - 1) & 7;
        if (f[(size_t)(y0 + DY[s]) * fw + x0 + DX[s]] != 0) { found = 1; break; }
    } while (s != s_end);
    if (!found) {   /* isolated pixel */
        f[(size_t)y0 * fw + x0] = -nbd;
        bl_push_pt(L, x0 - 1, y0 - 1);
        return;
    }
    const int x1 = x0 + DX[s], y1 = y0 + DY[s];
    int cx = x0, cy = y0;
    for (;;) {
        int right_zero_examined = 0, nx, ny;
        for (;;) {
            s = (s + 1) & 7;
            nx = cx + DX[s]; ny = cy + DY[s];
            if (f[(size_t)ny * fw + nx] != 0) break;
            if (s == 0) right_zero_examined = 1;
        }
        int32_t *c = &f[(size_t)cy * fw + cx];
        if (right_zero_examined) *c = -nbd;
        else if (*c == 1) *c = nbd;
        bl_push_pt(L, cx - 1, cy - 1);
        if (nx == x0 && ny == y0 && cx == x1 && cy == y1) break;
        cx = nx; cy = ny;
        s = (s + 4) & 7;
    }
}

/* Borders of a binary image (non-zero = foreground) in the dependency's RETR_LIST order (reverse discovery order).
 * pts: x,y pairs; offsets[i]..offsets[i+1] delimit border i.  Returns the number of borders, or -1 when a
 * capacity is exceeded. */
int orc_find_contours(const uint8_t *bin, int w, int h, int32_t *pts, long long max_pts, int64_t *offsets, int max_contours)
{
    const int fw = w + 2, fh = h + 2;
    int32_t *f = (int32_t *)calloc((size_t)fw * fh, sizeof(int32_t));
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) f[(size_t)(y + 1) * fw + x + 1] = bin[(size_t)y * w + x] ? 1 : 0;
    border_list L;
    memset(&L, 0, sizeof L);
    int nbd = 1;
    for (int y = 1; y <= h; y++) {
        for (int x = 1; x <= w; x++) {
            int32_t p = f[(size_t)y * fw + x];
            if (p == 0) continue;
            int outer = (p == 1 && f[(size_t)y * fw + x - 1] == 0);
            int hole = (!outer && p >= 1 && f[(size_t)y * fw + x + 1] == 0);
            if (!outer && !hole) continue;
            nbd++;
            bl_begin(&L);
            trace_border(f, fw, x, y, hole, nbd, &L);
        }
    }
    int n = L.n, rc = n;
    if (n > max_contours || (long long)L.n_pts > max_pts) rc = -1;
    else {
        if (L.n) L.start[L.n] = L.n_pts;
        size_t o = 0;
        for (int i = n - 1, k = 0; i >= 0; i--, k++) {
            size_t a = L.start[i], b = (i == n - 1) ? L.n_pts : L.start[i + 1];
            offsets[k] = (int64_t)o;
            memcpy(pts + 2 * o, L.pts + 2 * a, (b - a) * 2 * sizeof(int32_t));
            o += b - a;
        }
        offsets[n] = (int64_t)o;
    }
    free(L.pts); free(L.start); free(f);
    return rc;
}

/* ------------------------------------------------------------------------------------------------------ */
/* a6.C3: closed-curve Douglas-Peucker of the dependency (Appendix A.6) on integer points */
typedef struct { int start, end; } slice_t;

static double seg_dist2(const int32_t *a, const int32_t *b, const int32_t *p, double dx, double dy, double len2)
{
    double px = p[0] - a[0], py = p[1] - a[1];
    double proj = px * dx + py * dy;
    if (proj < 0) return px * px + py * py;
    if (proj > len2) { double qx = p[0] - b[0], qy = p[1] - b[1]; return qx * qx + qy * qy; }
    double cr = py * dx - px * dy;
    return cr * cr / len2;
}

int orc_approx_poly_dp(const int32_t *src, int count, double eps, int32_t *dst, int max_out)
{
    if (count == 0) return 0;
    int new_count = 0;
    eps *= eps;
    slice_t *stack = (slice_t *)malloc(((size_t)count + 8) * sizeof(slice_t));
    int top = 0;
    slice_t slice = {0, 0}, right = {0, 0};
    int pos = 0, le_eps = 0;
    const int32_t *start_pt = src;
#define WRITE_PT(p) do { if (new_count < max_out) { dst[2 * new_count] = (p)[0]; dst[2 * new_count + 1] = (p)[1]; } new_count++; } while (0)
    /* farthest point from the current start, three rounds */
    for (int it = 0; it < 3; it++) {
        double max_dist = 0;
        pos = (pos + right.start) % count;
        start_pt = src + 2 * pos;
        if (++pos >= count) pos = 0;
        for (int j = 1; j < count; j++) {
            const int32_t *pt = src + 2 * pos;
            if (++pos >= count) pos = 0;
            double dx = pt[0] - start_pt[0], dy = pt[1] - start_pt[1];
            double dist = dx * dx + dy * dy;
            if (dist > max_dist) { max_dist = dist; right.start = j; }
        }
        le_eps = max_dist <= eps;
    }
    if (!le_eps) {
        right.end = slice.start = pos % count;
        slice.end = right.start = (right.start + slice.start) % count;
        stack[top++] = right;
        stack[top++] = slice;
    } else {
        WRITE_PT(start_pt);
    }
    while (top > 0) {
        slice = stack[--top];
        const int32_t *end_pt = src + 2 * slice.end;
        pos = slice.start;
        start_pt = src + 2 * pos;
        if (++pos >= count) pos = 0;
        if (pos != slice.end) {
            double dx = end_pt[0] - start_pt[0], dy = end_pt[1] - start_pt[1];
            double len2 = dx * dx + dy * dy, max_dist = 0;
            while (pos != slice.end) {
                const int32_t *pt = src + 2 * pos;
                if (++pos >= count) pos = 0;
                double dist = seg_dist2(start_pt, end_pt, pt, dx, dy, len2);
                if (dist > max_dist) { max_dist = dist; right.start = (pos + count - 1) % count; }
            }
            le_eps = max_dist <= eps;
        } else {
            le_eps = 1;
            start_pt = src + 2 * slice.start;
        }
        if (le_eps) {
            WRITE_PT(start_pt);
        } else {
            right.end = slice.end;
            slice.end = right.start;
            stack[top++] = right;
            stack[top++] = slice;
        }
    }
    free(stack);
    if (new_count > max_out) return -new_count;   /* caller's buffer too small */
    /* clean-up: drop the middle vertex of nearly straight triples */
    int cnt = new_count;
    if (cnt == 0) return 0;
    int rp = cnt - 1, wpos;
    int32_t s[2], p[2], e[2];
#define READ_DST(v) do { (v)[0] = dst[2 * rp]; (v)[1] = dst[2 * rp + 1]; if (++rp >= cnt) rp = 0; } while (0)
    READ_DST(s);
    wpos = rp;
    READ_DST(p);
    for (int i = 0; i < cnt && new_count > 2; i++) {
        READ_DST(e);
        double dx = e[0] - s[0], dy = e[1] - s[1];
        double dist = fabs((double)(p[0] - s[0]) * dy - (double)(p[1] - s[1]) * dx);
        double sip = (double)(p[0] - s[0]) * (e[0] - p[0]) + (double)(p[1] - s[1]) * (e[1] - p[1]);
        if (dist * dist <= 0.5 * eps * (dx * dx + dy * dy) && dx != 0 && dy != 0 && sip >= 0) {
            new_count--;
            dst[2 * wpos] = s[0] = e[0]; dst[2 * wpos + 1] = s[1] = e[1];
            if (++wpos >= cnt) wpos = 0;
            READ_DST(p);
            i++;
            continue;
        }
        dst[2 * wpos] = s[0] = p[0]; dst[2 * wpos + 1] = s[1] = p[1];
        if (++wpos >= cnt) wpos = 0;
        p[0] = e[0]; p[1] = e[1];
    }
    return new_count;
#undef READ_DST
#undef WRITE_PT
}

/* Appendix A.6b */
int orc_is_contour_convex(const int32_t *p, int n)
{
    if (n < 3) return 0;   /* the dependency treats degenerate polygons as not usable here */
    int prev[2] = {p[2 * (n - 2)], p[2 * (n - 2) + 1]}, cur[2] = {p[2 * (n - 1)], p[2 * (n - 1) + 1]};
    long long dx0 = cur[0] - prev[0], dy0 = cur[1] - prev[1];
    int orientation = 0;
    for (int i = 0; i < n; i++) {
        prev[0] = cur[0]; prev[1] = cur[1];
        cur[0] = p[2 * i]; cur[1] = p[2 * i + 1];
        long long dx = cur[0] - prev[0], dy = cur[1] - prev[1];
        long long dxdy0 = dx * dy0, dydx0 = dy * dx0;
        orientation |= (dydx0 > dxdy0) ? 1 : ((dydx0 < dxdy0) ? 2 : 3);
        if (orientation == 3) return 0;
        dx0 = dx; dy0 = dy;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------------ */
/* CORNER_REFINE_CONTOUR (dependency: _refineCandidateLines / _interpolate2Dline / _getCrossPoint): the contour points are
 * grouped by the corner they follow in contour order (points ahead of the first corner join the last group), one least-squares
 * line per group (y = a x + b or x = a y + b, whichever extent is larger), refined corner = intersection of the two lines that
 * meet at it.  Bit-exact with cv2 4.13: sums in double, the 2 x 2 normal equations and their LU in float32, the intersection with
 * Matx22f::solve's closed form in float32.  Returns 0, or -1 when a corner is not a contour point. */
static void interp_line(double n, double sx, double sy, double sxx, double syy, double sxy, double minx, double maxx, double miny,
                        double maxy, float L[3])
{
    /* normal equations [[saa, sa], [sa, n]] (a, b)^T = (sab, sb): the sums are exact (the dependency accumulates the products of
     * its float32 matrices in double), stored as float32, then its float32 LU with partial pivoting */
    const int horiz = (float)maxx - (float)minx > (float)maxy - (float)miny;
    float A00 = (float)(horiz ? sxx : syy), A01 = (float)(horiz ? sx : sy), A10 = A01, A11 = (float)n;
    float B0 = (float)sxy, B1 = (float)(horiz ? sy : sx);
    if (fabsf(A10) > fabsf(A00)) { float t = A00; A00 = A10; A10 = t; t = A01; A01 = A11; A11 = t; t = B0; B0 = B1; B1 = t; }
    float x0 = 0, x1 = 0;
    if (!(fabsf(A00) < FLT_EPSILON)) {
        const float d = -1.f / A00, alpha = A10 * d;
        A11 = A11 + alpha * A01;
        B1 = B1 + alpha * B0;
        if (!(fabsf(A11) < FLT_EPSILON)) {
            x1 = B1 / A11;
            x0 = (B0 - A01 * x1) / A00;
        }
    }
    if (horiz) { L[0] = x0; L[1] = -1.f; L[2] = x1; }   /* y = a x + b */
    else { L[0] = -1.f; L[1] = x0; L[2] = x1; }         /* x = a y + b */
}

int orc_refine_candidate_lines(const int32_t *cont, int n, const float *corners, float *out)
{
    int idx[4] = {-1, -1, -1, -1};
    double S[5][10];
    for (int g = 0; g < 5; g++) { for (int k = 0; k < 6; k++) S[g][k] = 0; S[g][6] = S[g][8] = 1e300; S[g][7] = S[g][9] = -1e300; }
    int group = 4;
    for (int i = 0; i < n; i++) {
        double x = cont[2 * i], y = cont[2 * i + 1];
        for (int j = 0; j < 4; j++)
            if ((double)corners[2 * j] == x && (double)corners[2 * j + 1] == y) { idx[j] = i; group = j; }
        double *s = S[group];
        s[0] += 1; s[1] += x; s[2] += y; s[3] += x * x; s[4] += y * y; s[5] += x * y;
        if (x < s[6]) s[6] = x;
        if (x > s[7]) s[7] = x;
        if (y < s[8]) s[8] = y;
        if (y > s[9]) s[9] = y;
    }
    for (int j = 0; j < 4; j++) if (idx[j] < 0) return -1;
    if (S[4][0] > 0) {   /* points ahead of the first corner belong to the group that was open at the end */
        double *s = S[group], *e = S[4];
        for (int k = 0; k < 6; k++) s[k] += e[k];
        if (e[6] < s[6]) s[6] = e[6];
        if (e[7] > s[7]) s[7] = e[7];
        if (e[8] < s[8]) s[8] = e[8];
        if (e[9] > s[9]) s[9] = e[9];
    }
    int inc = 1;
    if (idx[0] > idx[1] && idx[3] > idx[0]) inc = -1;
    if (idx[2] > idx[3] && idx[1] > idx[2]) inc = -1;
    float L[4][3];
    for (int g = 0; g < 4; g++) interp_line(S[g][0], S[g][1], S[g][2], S[g][3], S[g][4], S[g][5], S[g][6], S[g][7], S[g][8], S[g][9], L[g]);
    for (int i = 0; i < 4; i++) {
        const float *a = L[i], *b = inc < 0 ? L[(i + 1) % 4] : L[(i + 3) % 4];
        const float b0 = -a[2], b1 = -b[2];
        const float det = a[0] * b[1] - a[1] * b[0];
        if (det == 0) { out[2 * i] = 0; out[2 * i + 1] = 0; continue; }   /* Matx::solve fails: the zero vector comes back */
        const float dinv = 1.f / det;
        out[2 * i] = (b0 * b[1] - b1 * a[1]) * dinv;
        out[2 * i + 1] = (b1 * a[0] - b0 * b[0]) * dinv;
    }
    return 0;
}

/* quads of one thresholded window in the dependency's order; returns count (may exceed max_quads) */
static int window_quads(const uint8_t *bin, int w, int h, const orc_classic_params *P, float *quads, float *refined, int have, int max_quads)
{
    long long max_pts = (long long)w * h * 2 + 16;
    int max_c = w * h / 2 + 16;
    int32_t *pts = (int32_t *)malloc((size_t)max_pts * 2 * sizeof(int32_t));
    int64_t *off = (int64_t *)malloc(((size_t)max_c + 1) * sizeof(int64_t));
    int n = orc_find_contours(bin, w, h, pts, max_pts, off, max_c);
    int mx = w > h ? w : h;
    unsigned min_px = (unsigned)(P->min_perimeter_rate * mx), max_px = (unsigned)(P->max_perimeter_rate * mx);
    int32_t *poly = (int32_t *)malloc(((size_t)max_px + 16) * 2 * sizeof(int32_t));
    for (int i = 0; i < n; i++) {
        unsigned cnt = (unsigned)(off[i + 1] - off[i]);
        if (cnt < min_px || cnt > max_px) continue;
        int m = orc_approx_poly_dp(pts + 2 * off[i], (int)cnt, (double)cnt * P->approx_accuracy_rate, poly, (int)max_px + 16);
        if (m != 4 || !orc_is_contour_convex(poly, 4)) continue;
        double min_d2 = (double)mx * mx;
        for (int j = 0; j < 4; j++) {
            int k = (j + 1) % 4;
            double d = (double)(poly[2 * j] - poly[2 * k]) * (poly[2 * j] - poly[2 * k]) +
                       (double)(poly[2 * j + 1] - poly[2 * k + 1]) * (poly[2 * j + 1] - poly[2 * k + 1]);
            if (d < min_d2) min_d2 = d;
        }
        double min_corner = (double)cnt * P->min_corner_distance_rate;
        if (min_d2 < min_corner * min_corner) continue;
        /* (4.13 applies minDistanceToBorder later, after the too-close grouping: oracle_decode.c) */
        if (have < max_quads) {
            float *q = quads + (size_t)have * 8;
            for (int j = 0; j < 8; j++) q[j] = (float)poly[j];
            /* clockwise order: swap corners 1 and 3 when the cross product is negative */
            double dx1 = q[2] - q[0], dy1 = q[3] - q[1], dx2 = q[4] - q[0], dy2 = q[5] - q[1];
            if (dx1 * dy2 - dy1 * dx2 < 0.0) {
                float tx = q[2], ty = q[3];
                q[2] = q[6]; q[3] = q[7]; q[6] = tx; q[7] = ty;
            }
            if (refined && orc_refine_candidate_lines(pts + 2 * off[i], (int)cnt, q, refined + (size_t)have * 8) != 0)
                memcpy(refined + (size_t)have * 8, q, 8 * sizeof(float));
        }
        have++;
    }
    free(poly); free(off); free(pts);
    return have;
}

/* a6.C1-C3: candidate quads of all windows (ascending window size), [n][8] float32.  Returns n (> max: overflow) */
int orc_classic_quads_refined(const uint8_t *gray, int w, int h, const orc_classic_params *P, float *quads, float *refined, int max_quads);
int orc_classic_quads(const uint8_t *gray, int w, int h, const orc_classic_params *P, float *quads, int max_quads)
{
    return orc_classic_quads_refined(gray, w, h, P, quads, NULL, max_quads);
}

/* as orc_classic_quads; refined (nullable) receives the CORNER_REFINE_CONTOUR corners of every candidate, same corner order */
int orc_classic_quads_refined(const uint8_t *gray, int w, int h, const orc_classic_params *P, float *quads, float *refined, int max_quads)
{
    int n_scales = (P->win_max - P->win_min) / P->win_step + 1;
    uint8_t *bin = (uint8_t *)malloc((size_t)w * h);
    int have = 0;
    for (int i = 0; i < n_scales; i++) {
        orc_adaptive_threshold(gray, w, h, P->win_min + i * P->win_step, P->constant, bin);
        have = window_quads(bin, w, h, P, quads, refined, have, max_quads);
    }
    free(bin);
    return have;
}

/* ------------------------------------------------------------------------------------------------------ */
/* a6.C4: cornerSubPix (Appendix A.7) */
static void rect_subpix_u8_f32(const uint8_t *im, int w, int h, float cx, float cy, int pw, int ph, float *dst)
{
    cx -= (pw - 1) * 0.5f;
    cy -= (ph - 1) * 0.5f;
    int ipx = (int)floorf(cx), ipy = (int)floorf(cy);
    float a = cx - ipx, b = cy - ipy;
    float a11 = (1.f - a) * (1.f - b), a12 = a * (1.f - b), a21 = (1.f - a) * b, a22 = a * b;
    if (ipx >= 0 && ipx + pw < w && ipy >= 0 && ipy + ph < h) {
        for (int i = 0; i < ph; i++) {
            const uint8_t *s = im + (size_t)(ipy + i) * w + ipx;
            for (int j = 0; j < pw; j++)
                dst[i * pw + j] = s[j] * a11 + s[j + 1] * a12 + s[j + w] * a21 + s[j + w + 1] * a22;
        }
        return;
    }
    /* patch crosses the image border: replicate the border pixels */
    for (int i = 0; i < ph; i++) {
        int y0 = ipy + i, y1 = y0 + 1;
        y0 = y0 < 0 ? 0 : y0 >= h ? h - 1 : y0;
        y1 = y1 < 0 ? 0 : y1 >= h ? h - 1 : y1;
        for (int j = 0; j < pw; j++) {
            int x0 = ipx + j, x1 = x0 + 1;
            x0 = x0 < 0 ? 0 : x0 >= w ? w - 1 : x0;
            x1 = x1 < 0 ? 0 : x1 >= w ? w - 1 : x1;
            dst[i * pw + j] = im[(size_t)y0 * w + x0] * a11 + im[(size_t)y0 * w + x1] * a12 + im[(size_t)y1 * w + x0] * a21 +
                              im[(size_t)y1 * w + x1] * a22;
        }
    }
}

void orc_corner_subpix(const uint8_t *im, int w, int h, float *corners, int n, int win, int max_iter, double eps)
{
    const int ww = 2 * win + 1;
    float *mask = (float *)malloc((size_t)ww * ww * sizeof(float));
    float *mx = (float *)malloc((size_t)ww * sizeof(float));
    float *patch = (float *)malloc((size_t)(ww + 2) * (ww + 2) * sizeof(float));
    for (int i = 0; i < ww; i++) {
        float x = (float)(i - win) / win;
        mx[i] = (float)exp(-x * x);
    }
    for (int i = 0; i < ww; i++)
        for (int j = 0; j < ww; j++) mask[i * ww + j] = mx[j] * mx[i];
    eps *= eps;
    for (int k = 0; k < n; k++) {
        float ctx = corners[2 * k], cty = corners[2 * k + 1], cix = ctx, ciy = cty;
        int iter = 0;
        double err = 0;
        do {
            rect_subpix_u8_f32(im, w, h, cix, ciy, ww + 2, ww + 2, patch);
            double a = 0, b = 0, c = 0, bb1 = 0, bb2 = 0;
            const int pp = ww + 2;
            for (int i = 0; i < ww; i++) {
                double py = i - win;
                for (int j = 0; j < ww; j++) {
                    const float *s = patch + (i + 1) * pp + (j + 1);
                    double m = mask[i * ww + j];
                    double tgx = s[1] - s[-1];
                    double tgy = s[pp] - s[-pp];
                    double gxx = tgx * tgx * m, gxy = tgx * tgy * m, gyy = tgy * tgy * m;
                    double px = j - win;
                    a += gxx; b += gxy; c += gyy;
                    bb1 += gxx * px + gxy * py;
                    bb2 += gxy * px + gyy * py;
                }
            }
            double det = a * c - b * b;
            if (fabs(det) <= DBL_EPSILON * DBL_EPSILON) break;
            double scale = 1.0 / det;
            float nx = (float)(cix + c * scale * bb1 - b * scale * bb2);
            float ny = (float)(ciy - b * scale * bb1 + a * scale * bb2);
            err = (double)(nx - cix) * (nx - cix) + (double)(ny - ciy) * (ny - ciy);
            cix = nx; ciy = ny;
            if (cix < 0 || cix >= w || ciy < 0 || ciy >= h) break;
        } while (++iter < max_iter && err > eps);
        if (fabsf(cix - ctx) > win || fabsf(ciy - cty) > win) { cix = ctx; ciy = cty; }
        corners[2 * k] = cix;
        corners[2 * k + 1] = ciy;
    }
    free(patch); free(mx); free(mask);
}
