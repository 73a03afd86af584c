/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the AprilTag-style quad detector that
 * cv2.aruco runs when cornerRefinementMethod == CORNER_REFINE_APRILTAG, which is what the reference
 * selects at aruco_detect.py:266 before calling aruco.detectMarkers at aruco_detect.py:267.
 *
 * The algorithm lives in OpenCV (un-vendored dependency, reference README.md:42).  Restated from the
 * published AprilTag-2 quad detector (Wang & Olson, IROS 2016) as specified in SURVEY.md Appendix A.3 and
 * pinned against the cv2 4.13.0 binary by tests/test_oracle_detect.py.  Not product code.
 *
 * Stages:
 *   A1 4x4-tile min/max ternary threshold     A2 union-find connected components
 *   A3 black/white boundary point clusters    A4 per-cluster quad fit
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

typedef struct {
    int min_cluster_pixels;
    int max_nmaxima;
    float critical_rad;
    float max_line_fit_mse;
    int min_white_black_diff;
} orc_at_params;

/* ------------------------------------------------------------------------------------------------------ */
/* fastAtan2 of the dependency: float32 polynomial, degrees (bit-exact vs cv2.fastAtan2, see tests) */
float orc_fast_atan2(float y, float x)
{
    const float k = (float)(180 / M_PI);
    const float p1 = 0.9997878412794807f * k, p3 = -0.3258083974640975f * k;
    const float p5 = 0.1555786518463281f * k, p7 = -0.04432655554792128f * k;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* ------------------------------------------------------------------------------------------------------ */
/* A1: tile min/max -> 3x3 tile dilation -> ternary {0,127,255} */
void orc_at_threshold(const uint8_t *im, int w, int h, int min_wb_diff, uint8_t *out)
{
    const int ts = 4;
    int tw = w / ts, th = h / ts;
    uint8_t *tmax = (uint8_t *)malloc((size_t)tw * th), *tmin = (uint8_t *)malloc((size_t)tw * th);
    uint8_t *dmax = (uint8_t *)malloc((size_t)tw * th), *dmin = (uint8_t *)malloc((size_t)tw * th);
    for (int ty = 0; ty < th; ty++)
        for (int tx = 0; tx < tw; tx++) {
            int mx = 0, mn = 255;
            for (int dy = 0; dy < ts; dy++)
                for (int dx = 0; dx < ts; dx++) {
                    int v = im[(size_t)(ty * ts + dy) * w + tx * ts + dx];
                    if (v > mx) mx = v;
                    if (v < mn) mn = v;
                }
            tmax[ty * tw + tx] = (uint8_t)mx;
            tmin[ty * tw + tx] = (uint8_t)mn;
        }
    for (int ty = 0; ty < th; ty++)
        for (int tx = 0; tx < tw; tx++) {
            int mx = 0, mn = 255;
            for (int dy = -1; dy <= 1; dy++) {
                if (ty + dy < 0 || ty + dy >= th) continue;
                for (int dx = -1; dx <= 1; dx++) {
                    if (tx + dx < 0 || tx + dx >= tw) continue;
                    int a = tmax[(ty + dy) * tw + tx + dx], b = tmin[(ty + dy) * tw + tx + dx];
                    if (a > mx) mx = a;
                    if (b < mn) mn = b;
                }
            }
            dmax[ty * tw + tx] = (uint8_t)mx;
            dmin[ty * tw + tx] = (uint8_t)mn;
        }
    memset(out, 0, (size_t)w * h);
    for (int ty = 0; ty < th; ty++)
        for (int tx = 0; tx < tw; tx++) {
            int mn = dmin[ty * tw + tx], mx = dmax[ty * tw + tx];
            int low = (mx - mn) < min_wb_diff;
            int thr = mn + (mx - mn) / 2;
            for (int dy = 0; dy < ts; dy++)
                for (int dx = 0; dx < ts; dx++) {
                    size_t o = (size_t)(ty * ts + dy) * w + tx * ts + dx;
                    out[o] = low ? 127 : (im[o] > thr ? 255 : 0);
                }
        }
    /* partial tiles at the right / bottom edge use the nearest full tile and are never marked 127 */
    for (int y = 0; y < h; y++) {
        int x0 = (y >= th * ts) ? 0 : tw * ts;
        int ty = y / ts;
        if (ty >= th) ty = th - 1;
        for (int x = x0; x < w; x++) {
            int tx = x / ts;
            if (tx >= tw) tx = tw - 1;
            int mx = dmax[ty * tw + tx], mn = dmin[ty * tw + tx];
            int thr = mn + (mx - mn) / 2;
            out[(size_t)y * w + x] = im[(size_t)y * w + x] > thr ? 255 : 0;
        }
    }
    free(tmax); free(tmin); free(dmax); free(dmin);
}

/* ------------------------------------------------------------------------------------------------------ */
/* A2: union-find; white 8-connected, black 4-connected, 127 skipped */
static uint32_t uf_find(uint32_t *p, uint32_t i)
{
    uint32_t r = i;
    while (p[r] != r) r = p[r];
    while (p[i] != r) { uint32_t n = p[i]; p[i] = r; i = n; }
    return r;
}
static void uf_union(uint32_t *p, uint32_t a, uint32_t b)
{
    a = uf_find(p, a); b = uf_find(p, b);
    if (a == b) return;
    if (a < b) p[b] = a; else p[a] = b; /* smaller index wins: root = raster-first pixel */
}

void orc_at_unionfind(const uint8_t *t, int w, int h, uint32_t *rep)
{
    size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; i++) rep[i] = (uint32_t)i;
    for (int y = 0; y < h - 1; y++)
        for (int x = 1; x < w - 1; x++) {
            uint8_t v = t[(size_t)y * w + x];
            if (v == 127) continue;
            uint32_t o = (uint32_t)(y * w + x);
            if (t[o + 1] == v) uf_union(rep, o, o + 1);
            if (t[o + w] == v) uf_union(rep, o, o + w);
            if (v == 255) {
                if (t[o + w - 1] == v) uf_union(rep, o, o + w - 1);
                if (t[o + w + 1] == v) uf_union(rep, o, o + w + 1);
            }
        }
    for (size_t i = 0; i < n; i++) rep[i] = uf_find(rep, (uint32_t)i);
}

/* ------------------------------------------------------------------------------------------------------ */
/* A3: boundary points keyed by the unordered pair of component representatives */
typedef struct {
    uint64_t key;
    uint32_t seq; /* raster emission order (keeps the dependency's insertion order inside a cluster) */
    uint16_t x, y;
    int16_t gx, gy;
    float theta;
} orc_pt;

static int cmp_key_seq(const void *a, const void *b)
{
    const orc_pt *p = (const orc_pt *)a, *q = (const orc_pt *)b;
    if (p->key != q->key) return p->key < q->key ? -1 : 1;
    return p->seq < q->seq ? -1 : p->seq > q->seq;
}
static int cmp_theta_seq(const void *a, const void *b)
{
    const orc_pt *p = (const orc_pt *)a, *q = (const orc_pt *)b;
    if (p->theta != q->theta) return p->theta < q->theta ? -1 : 1;
    return p->seq < q->seq ? -1 : p->seq > q->seq;
}

static size_t emit_points(const uint8_t *t, const uint32_t *rep, int w, int h, orc_pt **out)
{
    size_t cap = 1 << 16, n = 0;
    orc_pt *pts = (orc_pt *)malloc(cap * sizeof(orc_pt));
    static const int DX[4] = {1, 0, -1, 1}, DY[4] = {0, 1, 1, 1};
    for (int y = 1; y < h - 1; y++)
        for (int x = 1; x < w - 1; x++) {
            int v0 = t[(size_t)y * w + x];
            if (v0 == 127) continue;
            uint64_t rep0 = rep[(size_t)y * w + x];
            for (int k = 0; k < 4; k++) {
                int dx = DX[k], dy = DY[k];
                int v1 = t[(size_t)(y + dy) * w + x + dx];
                if (v0 + v1 != 255) continue;
                uint64_t rep1 = rep[(size_t)(y + dy) * w + x + dx];
                if (n == cap) { cap *= 2; pts = (orc_pt *)realloc(pts, cap * sizeof(orc_pt)); }
                orc_pt *p = &pts[n];
                p->key = rep0 < rep1 ? (rep1 << 32) + rep0 : (rep0 << 32) + rep1;
                p->seq = (uint32_t)n;
                p->x = (uint16_t)(2 * x + dx);
                p->y = (uint16_t)(2 * y + dy);
                p->gx = (int16_t)(dx * (v1 - v0));
                p->gy = (int16_t)(dy * (v1 - v0));
                p->theta = 0;
                n++;
            }
        }
    *out = pts;
    return n;
}

/* ------------------------------------------------------------------------------------------------------ */
/* A4: quad fit */
typedef struct { double Mx, My, Mxx, Mxy, Myy, W; } lfp;

static void fit_line(const lfp *l, int sz, int i0, int i1, double *lineparm, double *err, double *mse)
{
    double Mx, My, Mxx, Mxy, Myy, W;
    int N;
    if (i0 < i1) {
        N = i1 - i0 + 1;
        Mx = l[i1].Mx; My = l[i1].My; Mxx = l[i1].Mxx; Mxy = l[i1].Mxy; Myy = l[i1].Myy; W = l[i1].W;
        if (i0 > 0) {
            Mx -= l[i0 - 1].Mx; My -= l[i0 - 1].My; Mxx -= l[i0 - 1].Mxx;
            Mxy -= l[i0 - 1].Mxy; Myy -= l[i0 - 1].Myy; W -= l[i0 - 1].W;
        }
    } else {
        Mx = l[sz - 1].Mx - l[i0 - 1].Mx; My = l[sz - 1].My - l[i0 - 1].My;
        Mxx = l[sz - 1].Mxx - l[i0 - 1].Mxx; Mxy = l[sz - 1].Mxy - l[i0 - 1].Mxy;
        Myy = l[sz - 1].Myy - l[i0 - 1].Myy; W = l[sz - 1].W - l[i0 - 1].W;
        Mx += l[i1].Mx; My += l[i1].My; Mxx += l[i1].Mxx; Mxy += l[i1].Mxy; Myy += l[i1].Myy; W += l[i1].W;
        N = sz - i0 + i1 + 1;
    }
    double Ex = Mx / W, Ey = My / W;
    double Cxx = Mxx / W - Ex * Ex, Cxy = Mxy / W - Ex * Ey, Cyy = Myy / W - Ey * Ey;
    float normal_theta = (float)(.5f * (M_PI / 180)) * orc_fast_atan2((float)(-2 * Cxy), (float)(Cyy - Cxx));
    double nx = cosf(normal_theta), ny = sinf(normal_theta);
    if (lineparm) { lineparm[0] = Ex; lineparm[1] = Ey; lineparm[2] = nx; lineparm[3] = ny; }
    if (err) *err = nx * nx * N * Cxx + 2 * nx * ny * N * Cxy + ny * ny * N * Cyy;
    if (mse) *mse = nx * nx * Cxx + 2 * nx * ny * Cxy + ny * ny * Cyy;
}

static int cmp_double_desc(const void *a, const void *b)
{
    double x = *(const double *)a, y = *(const double *)b;
    return x > y ? -1 : x < y;
}

static int segment_maxima(const orc_at_params *P, int sz, const lfp *l, int indices[4])
{
    int ksz = 20 < sz / 12 ? 20 : sz / 12;
    if (ksz < 2) return 0;
    double *errs = (double *)malloc(sizeof(double) * sz), *y = (double *)malloc(sizeof(double) * sz);
    for (int i = 0; i < sz; i++) fit_line(l, sz, (i + sz - ksz) % sz, (i + ksz) % sz, NULL, &errs[i], NULL);
    {
        double sigma = 1, cutoff = 0.05;
        int fsz = (int)floor(sqrt(-log(cutoff) * 2 * sigma * sigma)) + 1;
        fsz = 2 * fsz + 1;
        float f[32];
        for (int i = 0; i < fsz; i++) {
            int j = i - fsz / 2;
            f[i] = (float)exp(-j * j / (2 * sigma * sigma));
        }
        for (int iy = 0; iy < sz; iy++) {
            double acc = 0;
            for (int i = 0; i < fsz; i++) acc += errs[(iy + i - fsz / 2 + sz) % sz] * f[i];
            y[iy] = acc;
        }
        memcpy(errs, y, sizeof(double) * sz);
    }
    int *maxima = (int *)malloc(sizeof(int) * sz);
    double *maxima_errs = (double *)malloc(sizeof(double) * sz);
    int nmaxima = 0, ok = 0;
    for (int i = 0; i < sz; i++)
        if (errs[i] > errs[(i + 1) % sz] && errs[i] > errs[(i + sz - 1) % sz]) {
            maxima[nmaxima] = i;
            maxima_errs[nmaxima] = errs[i];
            nmaxima++;
        }
    if (nmaxima < 4) goto done;
    if (nmaxima > P->max_nmaxima) {
        double *copy = (double *)malloc(sizeof(double) * nmaxima);
        memcpy(copy, maxima_errs, sizeof(double) * nmaxima);
        qsort(copy, nmaxima, sizeof(double), cmp_double_desc);
        double thresh = copy[P->max_nmaxima];
        int out = 0;
        for (int in = 0; in < nmaxima; in++) {
            if (maxima_errs[in] <= thresh) continue;
            maxima[out++] = maxima[in];
        }
        nmaxima = out;
        free(copy);
    }
    {
        int best[4] = {0, 0, 0, 0};
        double best_error = HUGE_VALF;
        double err01, err12, err23, err30, mse01, mse12, mse23, mse30;
        double p01[4], p12[4], p23[4], p30[4];
        double max_dot = cos(P->critical_rad);
        for (int m0 = 0; m0 < nmaxima - 3; m0++) {
            int i0 = maxima[m0];
            for (int m1 = m0 + 1; m1 < nmaxima - 2; m1++) {
                int i1 = maxima[m1];
                fit_line(l, sz, i0, i1, p01, &err01, &mse01);
                if (mse01 > P->max_line_fit_mse) continue;
                for (int m2 = m1 + 1; m2 < nmaxima - 1; m2++) {
                    int i2 = maxima[m2];
                    fit_line(l, sz, i1, i2, p12, &err12, &mse12);
                    if (mse12 > P->max_line_fit_mse) continue;
                    double dot = p01[2] * p12[2] + p01[3] * p12[3];
                    if (fabs(dot) > max_dot) continue;
                    for (int m3 = m2 + 1; m3 < nmaxima; m3++) {
                        int i3 = maxima[m3];
                        fit_line(l, sz, i2, i3, p23, &err23, &mse23);
                        if (mse23 > P->max_line_fit_mse) continue;
                        fit_line(l, sz, i3, i0, p30, &err30, &mse30);
                        if (mse30 > P->max_line_fit_mse) continue;
                        double err = err01 + err12 + err23 + err30;
                        if (err < best_error) {
                            best_error = err;
                            best[0] = i0; best[1] = i1; best[2] = i2; best[3] = i3;
                        }
                    }
                }
            }
        }
        if (best_error == HUGE_VALF) goto done;
        for (int i = 0; i < 4; i++) indices[i] = best[i];
        if (best_error / sz < P->max_line_fit_mse) ok = 1;
    }
done:
    free(errs); free(y); free(maxima); free(maxima_errs);
    return ok;
}

static double sq(double v) { return v * v; }

/* pts[0..sz) is one cluster; returns 1 and fills quad[4][2] (p0..p3 of the fit) */
static int fit_quad(const orc_at_params *P, const uint8_t *im, int w, int h, orc_pt *pts, int sz, float quad[4][2])
{
    if (sz < 4) return 0;
    int xmax = 0, xmin = INT32_MAX, ymax = 0, ymin = INT32_MAX;
    for (int i = 0; i < sz; i++) {
        if (pts[i].x > xmax) xmax = pts[i].x;
        if (pts[i].x < xmin) xmin = pts[i].x;
        if (pts[i].y > ymax) ymax = pts[i].y;
        if (pts[i].y < ymin) ymin = pts[i].y;
    }
    double cx = (xmin + xmax) * 0.5 + 0.05118, cy = (ymin + ymax) * 0.5 + -0.028581;
    double dot = 0;
    for (int i = 0; i < sz; i++) {
        double dx = pts[i].x - cx, dy = pts[i].y - cy;
        pts[i].theta = orc_fast_atan2((float)dy, (float)dx) * (float)(M_PI / 180);
        dot += dx * pts[i].gx + dy * pts[i].gy;
    }
    if (dot < 0) return 0;
    qsort(pts, sz, sizeof(orc_pt), cmp_theta_seq);
    {
        int outpos = 1;
        int last = 0;
        for (int i = 1; i < sz; i++) {
            if (pts[i].x != pts[last].x || pts[i].y != pts[last].y) {
                if (i != outpos) pts[outpos] = pts[i];
                outpos++;
            }
            last = i;
        }
        /* NB: `last` indexes the pre-compaction slot i; slots >= outpos are still intact when compared,
           and slot i itself is only overwritten by later iterations, so this equals comparing with the
           previous input element as the dependency does. */
        sz = outpos;
    }
    if (sz < 4) return 0;
    lfp *l = (lfp *)calloc(sz, sizeof(lfp));
    for (int i = 0; i < sz; i++) {
        if (i > 0) l[i] = l[i - 1];
        double x = pts[i].x * .5 + 0.5, y = pts[i].y * .5 + 0.5;
        int ix = (int)x, iy = (int)y;
        double W = 1;
        if (ix > 0 && ix + 1 < w && iy > 0 && iy + 1 < h) {
            int gx = im[(size_t)iy * w + ix + 1] - im[(size_t)iy * w + ix - 1];
            int gy = im[(size_t)(iy + 1) * w + ix] - im[(size_t)(iy - 1) * w + ix];
            W = sqrt((double)(gx * gx + gy * gy)) + 1;
        }
        double fx = x, fy = y;
        l[i].Mx += W * fx;
        l[i].My += W * fy;
        l[i].Mxx += W * fx * fx;
        l[i].Mxy += W * fx * fy;
        l[i].Myy += W * fy * fy;
        l[i].W += W;
    }
    int res = 0, indices[4];
    double lines[4][4];
    if (!segment_maxima(P, sz, l, indices)) goto finish;
    for (int i = 0; i < 4; i++) {
        double mse;
        fit_line(l, sz, indices[i], indices[(i + 1) & 3], lines[i], NULL, &mse);
        if (mse > P->max_line_fit_mse) goto finish;
    }
    for (int i = 0; i < 4; i++) {
        int i1 = (i + 1) & 3;
        double A00 = lines[i][3], A01 = -lines[i1][3];
        double A10 = -lines[i][2], A11 = lines[i1][2];
        double B0 = -lines[i][0] + lines[i1][0];
        double B1 = -lines[i][1] + lines[i1][1];
        double det = A00 * A11 - A10 * A01;
        if (fabs(det) < 0.001) goto finish;
        double det_inv = 1.0 / det;
        double W00 = A11 * det_inv, W01 = -A01 * det_inv;
        double L0 = W00 * B0 + W01 * B1;
        quad[i][0] = (float)(lines[i][0] + L0 * A00);
        quad[i][1] = (float)(lines[i][1] + L0 * A10);
    }
    res = 1;
    {
        double area = 0, length[3], p;
        for (int i = 0; i < 3; i++) {
            int a = i, b = (i + 1) % 3;
            length[i] = sqrt(sq(quad[b][0] - quad[a][0]) + sq(quad[b][1] - quad[a][1]));
        }
        p = (length[0] + length[1] + length[2]) / 2;
        area += sqrt(p * (p - length[0]) * (p - length[1]) * (p - length[2]));
        static const int idxs[4] = {2, 3, 0, 2};
        for (int i = 0; i < 3; i++) {
            int a = idxs[i], b = idxs[i + 1];
            length[i] = sqrt(sq(quad[b][0] - quad[a][0]) + sq(quad[b][1] - quad[a][1]));
        }
        p = (length[0] + length[1] + length[2]) / 2;
        area += sqrt(p * (p - length[0]) * (p - length[1]) * (p - length[2]));
        int d = 8;
        if (area < d * d) { res = 0; goto finish; }
    }
    {
        double total = 0;
        for (int i = 0; i < 4; i++) {
            int i0 = i, i1 = (i + 1) & 3, i2 = (i + 2) & 3;
            double theta0 = atan2f(quad[i0][1] - quad[i1][1], quad[i0][0] - quad[i1][0]);
            double theta1 = atan2f(quad[i2][1] - quad[i1][1], quad[i2][0] - quad[i1][0]);
            double dtheta = theta0 - theta1;
            if (dtheta < 0) dtheta += 2 * M_PI;
            if (dtheta < P->critical_rad || dtheta > (M_PI - P->critical_rad)) res = 0;
            total += dtheta;
        }
        if (total < 6.2 || total > 6.4) { res = 0; goto finish; }
    }
finish:
    free(l);
    return res;
}

/* Full quad detector.  quads_out: [max_quads][8] floats in the dependency's candidate order
 * (p3,p0,p1,p2).  stats[0]=#points, stats[1]=#clusters, stats[2]=#clusters passing the size filter.
 * Optional dumps: thresh_out (w*h), rep_out (w*h). */
int orc_at_quads(const uint8_t *im, int w, int h, const orc_at_params *P, float *quads_out, int max_quads,
                 int64_t *stats, uint8_t *thresh_out, uint32_t *rep_out)
{
    size_t n = (size_t)w * h;
    uint8_t *t = thresh_out ? thresh_out : (uint8_t *)malloc(n);
    uint32_t *rep = rep_out ? rep_out : (uint32_t *)malloc(n * sizeof(uint32_t));
    orc_at_threshold(im, w, h, P->min_white_black_diff, t);
    orc_at_unionfind(t, w, h, rep);
    orc_pt *pts;
    size_t np = emit_points(t, rep, w, h, &pts);
    qsort(pts, np, sizeof(orc_pt), cmp_key_seq);
    int nq = 0;
    int64_t nclusters = 0, nfit = 0;
    size_t i = 0;
    while (i < np) {
        size_t j = i;
        while (j < np && pts[j].key == pts[i].key) j++;
        size_t sz = j - i;
        nclusters++;
        if ((int64_t)sz >= P->min_cluster_pixels && sz <= (size_t)(3 * (2 * w + 2 * h))) {
            float q[4][2];
            nfit++;
            if (fit_quad(P, im, w, h, pts + i, (int)sz, q)) {
                if (nq < max_quads) {
                    static const int order[4] = {3, 0, 1, 2};
                    for (int k = 0; k < 4; k++) {
                        quads_out[nq * 8 + 2 * k] = q[order[k]][0];
                        quads_out[nq * 8 + 2 * k + 1] = q[order[k]][1];
                    }
                }
                nq++;
            }
        }
        i = j;
    }
    if (stats) { stats[0] = (int64_t)np; stats[1] = nclusters; stats[2] = nfit; }
    free(pts);
    if (!thresh_out) free(t);
    if (!rep_out) free(rep);
    return nq;
}
