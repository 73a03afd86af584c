// sequence.cu -- the per-sequence post-pass of aruco_detect.py:598-782 (marker gating, marker-length recurrence, LED
// read-out, ground-plane vehicle distances, CSV row of :146-185) as native code:
//   apse_sequence_scan    host, sequential: the O(markers) state machine over the frames in order (the only part of the
//                         pipeline that is order-dependent); every cv2.projectPoints call of the reference (:344,424,468)
//                         is NOT evaluated here but deferred as a job -- the projections feed outputs only, never the state
//   apse_sequence_jobs    device, one warp per job: LED strip read-out (:338-373) and nearest vehicle-outline point
//                         (:433-492) with the FP64 projection of pose.cu; independent across frames
//   apse_sequence_finish  host: job results -> rows (values the reference leaves stale between frames stay stale),
//                         Python round() semantics of :146-185
//   apse_sequence_csv     host: the text the reference writes (:131-139,146-185), str(float) formatting of Python
// apse_uav_b200/postpass.py is the readable host mirror of the same logic (used by the parity tests to cross-check this file).
#include "common.cuh"
#include <cmath>
#include <math.h>
#include <float.h>
#include <string.h>
#include <stdlib.h>
#include <charconv>
#include <vector>

// ---------------------------------------------------------------------------------------------------------
// constants of the reference (aruco_detect.py:542-549,583-586,340-341)
static const double VEH_DIM[5][4] = {{0, 0, 0, 0},
                                     {-1.95, 2.8, -0.9, 0.9},      // vehicle 1: back, front, left, right
                                     {-1.68, 2.86, -0.87, 0.87},   // vehicle 2
                                     {-1.32, 2.48, -0.86, 0.86},   // vehicle 3
                                     {-2.35, 2.49, -0.86, 0.86}};  // host (id 4)
__constant__ float c_led_axis[8][3] = {{-0.419f, -0.42f, 0}, {-0.414f, -0.305f, 0}, {-0.409f, -0.19f, 0}, {-0.404f, -0.07f, 0},
                                       {-0.399f, 0.065f, 0}, {-0.393f, 0.19f, 0},  {-0.388f, 0.315f, 0}, {-0.382f, 0.435f, 0}};

// Python's round(x, nd) on a float: the exact binary value of x rounded half-to-even to nd decimals (what _Py_dg_dtoa mode 3
// yields), converted back to the nearest double.  For the magnitudes of this path (|x| * 10^nd < 2^53, nd <= 5) that is exact
// 128-bit integer arithmetic: x = m * 2^e, q = round_half_even(m * 10^nd * 2^e), result = q / 10^nd (one IEEE division of two
// exactly representable integers = the correctly rounded value of the decimal string).  Anything else takes the string path.
static double py_round_slow(double x, int nd)
{
    char buf[400];
    snprintf(buf, sizeof buf, "%.*f", nd, x);
    return strtod(buf, nullptr);
}

static double py_round(double x, int nd)
{
    if (!isfinite(x) || x == 0) return x;
    static const double P10[6] = {1, 10, 100, 1000, 10000, 100000};
    const double a = fabs(x);
    if (nd < 0 || nd > 5 || a * P10[nd] >= 4.0e15 || a < 1e-300) return py_round_slow(x, nd);
    {
        // fast path: y = a * 10^nd carries a relative error of 2^-53, i.e. less than 1e-6 absolute below 2^32; away from a tie by
        // more than that the nearest integer of y is the nearest integer of the exact product, and the result is the same
        // correctly rounded quotient the exact path returns
        const double y = a * P10[nd];
        if (y < 4294967296.0) {
            const double fl = floor(y), fr = y - fl;
            if (fabs(fr - 0.5) > 1e-6) {
                const double r = (fr > 0.5 ? fl + 1.0 : fl) / P10[nd];
                return x < 0 ? -r : r;
            }
        }
    }
    int e;
    const double fr = frexp(a, &e);                                   // a = fr * 2^e, fr in [0.5, 1)
    const unsigned long long m = (unsigned long long)ldexp(fr, 53);   // 53-bit integer mantissa, a = m * 2^(e - 53)
    const int sh = 53 - e;                                            // a * 10^nd = (m * 10^nd) / 2^sh
    const unsigned __int128 prod = (unsigned __int128)m * (unsigned long long)P10[nd];
    unsigned long long q;
    if (sh <= 0) {
        q = (unsigned long long)(prod << (-sh));
    } else if (sh >= 127) {
        q = 0;
    } else {
        const unsigned __int128 one = (unsigned __int128)1 << sh, rem = prod & (one - 1), half = one >> 1;
        q = (unsigned long long)(prod >> sh);
        if (rem > half || (rem == half && (q & 1))) q++;
    }
    const double r = (double)q / P10[nd];
    return x < 0 ? -r : r;
}

struct SeqState {   // module globals of the reference that survive from frame to frame
    double marker_length;
    int detected_prev[4];
    double prev_xy[5][2];
    double cxy[5][2], diff[5];
    double msp[5], size_corr[5];
    bool have_msp[5];
    double altitude;
    // values apse_sequence_finish leaves stale between frames
    int leds;
    double dist[3][2];
    int initialised;
};
static_assert(sizeof(SeqState) <= sizeof(apse_seq_state), "apse_seq_state too small");

static double marker_length_correction(const apse_seq_config &c, double altitude)
{
    return c.marker_length_org * (1 - 0.00057 * altitude / c.marker_div) / c.div;   // :306-308, same evaluation order
}

// first angle of scipy's Rotation.from_rotvec(rvec).as_euler('zxy', degrees=True) (:412-413): the two matrix entries it is the
// arctangent of
static void yaw_zxy_terms(const double r[3], double &R10, double &R11)
{
    const double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (th < 1e-300) { R10 = 0.0; R11 = 1.0; return; }
    const double k[3] = {r[0] / th, r[1] / th, r[2] / th};
    const double Kx[9] = {0, -k[2], k[1], k[2], 0, -k[0], -k[1], k[0], 0};
    double K2[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) K2[i * 3 + j] = Kx[i * 3] * Kx[j] + Kx[i * 3 + 1] * Kx[3 + j] + Kx[i * 3 + 2] * Kx[6 + j];
    const double s = sin(th), c1 = 1 - cos(th);
    R10 = s * Kx[3] + c1 * K2[3];
    R11 = 1.0 + s * Kx[4] + c1 * K2[4];
}

// round(yaw, 2) < 0 (:413-414), the only use the path makes of the angle.  yaw = atan2(R10, R11) in degrees is negative exactly
// when R10 < 0, and rounds to a negative number unless |yaw| <= 0.005 degrees; only within a hundredfold margin of that
// boundary the angle itself is evaluated and rounded as Python does.
static bool yaw_rounds_negative(const double r[3])
{
    double R10, R11;
    yaw_zxy_terms(r, R10, R11);
    if (!(R10 < 0)) {
        if (R10 > 0 || R11 >= 0 || std::signbit(R10) == 0) return false;   // yaw >= +0
        return py_round(atan2(R10, R11) * (180.0 / M_PI), 2) < 0;           // R10 = -0, R11 < 0: atan2 gives -pi
    }
    if (R11 <= 0 || -R10 > 1e-2 * R11) return true;                         // |yaw| > 0.57 degrees
    return py_round(atan2(R10, R11) * (180.0 / M_PI), 2) < 0;
}

extern "C" {

double apse_py_round(double x, int ndigits) { return py_round(x, ndigits); }

void apse_seq_config_default(apse_seq_config *c)
{
    if (!c) return;
    c->start_frame = 1; c->step_frame = 1;
    c->marker_length_org = 0.55; c->marker_div = 1.2; c->div = 1.013;
    c->width = 3840; c->height = 2160;
    c->leds_threshold = -1;
    c->leds = 0;
}

int apse_sequence_scan(const apse_seq_config *cfg, int n_frames, int max_markers, const int32_t *n_markers, const int32_t *ids,
                       const float *corners, const double *rvec, const double *tvec, int rescale_tvec, double *lengths,
                       apse_seq_row *rows, apse_seq_job *jobs, int job_cap, int *n_jobs)
{
    apse_seq_state st;
    memset(&st, 0, sizeof st);
    return apse_sequence_scan_chunk(cfg, &st, 0, n_frames, max_markers, n_markers, ids, corners, rvec, tvec, rescale_tvec, lengths, rows, jobs,
                                    job_cap, n_jobs);
}

int apse_sequence_scan_chunk(const apse_seq_config *cfg, apse_seq_state *state, int frame0, int n_frames, int max_markers,
                             const int32_t *n_markers, const int32_t *ids, const float *corners, const double *rvec, const double *tvec,
                             int rescale_tvec, double *lengths, apse_seq_row *rows, apse_seq_job *jobs, int job_cap, int *n_jobs)
{
    if (!cfg || !state || frame0 < 0 || n_frames < 0 || max_markers <= 0) return APSE_ERR_INVALID_ARG;
    if (n_frames > 0 && (!n_markers || !ids || !corners || !rvec || !tvec)) return APSE_ERR_INVALID_ARG;
    if ((rows || jobs) && (!rows || !jobs || !n_jobs || job_cap < 0)) return APSE_ERR_INVALID_ARG;
    const apse_seq_config &c = *cfg;
    const double diff_max = 2.0 / 3 * c.step_frame * 2;                                    // :524
    SeqState &S = *reinterpret_cast<SeqState *>(state);
    if (!S.initialised) {
        memset(&S, 0, sizeof S);
        S.marker_length = c.marker_length_org;                                             // :521
        S.initialised = 1;
    }
    int nj = 0;
    std::vector<int32_t> idl((size_t)max_markers);
    for (int fr = 0; fr < n_frames; fr++) {
        const int k = c.start_frame + (frame0 + fr) * c.step_frame;
        const bool first = frame0 + fr == 0;
        if (lengths) lengths[fr] = S.marker_length;    // the marker length estimatePoseSingleMarkers sees in this frame (:601)
        const double tscale = rescale_tvec ? S.marker_length / c.marker_length_org : 1.0;   // tvec is linear in the marker length
        const int n = n_markers[fr] < max_markers ? n_markers[fr] : max_markers;
        int detected[4] = {0, 0, 0, 0};
        double dims[5][4];
        memcpy(dims, VEH_DIM, sizeof dims);                                                // re-initialised every frame (:583-586)
        apse_seq_row *row = rows ? rows + fr : nullptr;
        if (row) {
            memset(row, 0, sizeof *row);
            row->frame_id = k;
            row->job_led = -1;
            row->job_dist[0] = row->job_dist[1] = row->job_dist[2] = -1;
        }
        if (n > 0) {
            const int32_t *fid = ids + (size_t)fr * max_markers;
            const float *fc = corners + (size_t)fr * max_markers * 8;
            const double *frv = rvec + (size_t)fr * max_markers * 3, *ftv = tvec + (size_t)fr * max_markers * 3;
            for (int i = 0; i < n; i++) idl[i] = fid[i];
            for (int i = 0; i < n; i++) {
                const float *q = fc + 8 * i;
                const double tz = ftv[3 * i + 2] * tscale;
                // the reference tests id 4, then the "[4] not in ids" altitude fallback, then ids 1, 2, 3 (:606-723)
                static const int order[5] = {4, 0, 1, 2, 3};
                for (int oi = 0; oi < 5; oi++) {
                    const int vid = order[oi];
                    if (vid == 0) {
                        bool has4 = false;
                        for (int j = 0; j < n; j++) has4 |= idl[j] == 4;
                        if (!has4) {                                                       // :639-642
                            S.altitude = tz;
                            S.marker_length = marker_length_correction(c, S.altitude);
                            S.altitude = S.altitude / c.marker_div;
                        }
                        continue;
                    }
                    if (idl[i] != vid) continue;
                    const int slot = vid - 1;
                    // getMarkerData (:271-288): float32 corner sums truncated by int() before the division; float32 side lengths
                    const float sx = ((q[0] + q[2]) + q[4]) + q[6], sy = ((q[1] + q[3]) + q[5]) + q[7];
                    const double cx = fabs((double)(long long)sx / 4), cy = fabs((double)(long long)sy / 4);
                    auto side = [&](int a, int b) {
                        const float dx = q[2 * a] - q[2 * b], dy = q[2 * a + 1] - q[2 * b + 1];
                        return sqrtf(dx * dx + dy * dy);
                    };
                    const float msp = (((side(1, 0) + side(2, 1)) + side(3, 2)) + side(0, 3)) / 4;
                    double diff = 0;
                    if (!first) {
                        // the reference passes the UNSIGNED previous centre and compares with the signed current one; both are
                        // non-negative for corners inside the image
                        const double cxs = (double)(long long)sx / 4, cys = (double)(long long)sy / 4;
                        const double ddx = S.prev_xy[vid][0] - cxs, ddy = S.prev_xy[vid][1] - cys;
                        diff = sqrt(ddx * ddx + ddy * ddy) * S.marker_length / msp;
                    }
                    S.cxy[vid][0] = cx; S.cxy[vid][1] = cy; S.diff[vid] = diff;
                    if (S.detected_prev[slot] == 0) {                                      // new marker or false positive
                        detected[slot] = 1;
                        S.prev_xy[vid][0] = cx; S.prev_xy[vid][1] = cy;
                    }
                    if ((S.detected_prev[slot] == 1 && diff < diff_max) || first) {
                        detected[slot] = 1;
                        if (row && i < 31) row->accepted_mask |= 1 << i;
                        if (vid == 4) {                                                    // :622-624
                            S.altitude = tz;
                            S.marker_length = marker_length_correction(c, S.altitude);
                            S.altitude = S.altitude / c.marker_div;
                        }
                        // calculateAverageMarkerSize with N_avg = 1 (:290-304): np.count_nonzero returns a 64-bit numpy integer, so
                        // float32 msp * count is promoted to float64 and size_corr / the rescaled msp are float64 from here on
                        const double size_corr = (double)msp / ((double)msp * 1.0);
                        S.size_corr[vid] = size_corr;
                        S.msp[vid] = (double)msp * size_corr;
                        S.have_msp[vid] = true;
                        const double tv[3] = {ftv[3 * i] * tscale, ftv[3 * i + 1] * tscale, tz};
                        if (vid == 4 && c.leds && jobs) {                                  // detectAndDrawLEDs (:338-373), deferred
                            if (nj >= job_cap) return APSE_ERR_CAPACITY;
                            apse_seq_job &J = jobs[nj];
                            memset(&J, 0, sizeof J);
                            J.frame = frame0 + fr; J.kind = 0;
                            for (int a = 0; a < 3; a++) { J.rvec[a] = frv[3 * i + a]; J.tvec[a] = tv[a] / size_corr; }
                            int thr = c.leds_threshold;
                            if (thr < 0) { thr = 190 + (int)(tv[2] / c.marker_div); if (thr < 240) thr = 240; }   // max(190 + int(..), 240)
                            J.led_threshold = thr;
                            if (row) row->job_led = nj;
                            nj++;
                        }
                        S.prev_xy[vid][0] = cx; S.prev_xy[vid][1] = cy;
                        if (jobs && vid != 4) {   // drawBoundingBox (:406-420), the dimension scaling only; it feeds the distance jobs,
                                                  // which read the outline of vehicles 1-3 (the host's own outline is only drawn)
                            double ah = atan(tv[0] / tv[2]), av = atan(tv[1] / tv[2]);
                            if (!yaw_rounds_negative(frv + 3 * i)) { ah = -ah; av = -av; }
                            dims[vid][0] *= 1 - ah / 2; dims[vid][1] *= 1 + ah / 2;
                            dims[vid][2] *= 1 - av / 2; dims[vid][3] *= 1 + av / 2;
                        }
                    } else {
                        idl[i] = -1;                                                       // :637,669,696,723
                    }
                }
            }
            // distances from the host marker to every vehicle still carrying its id (:729-780), deferred
            if (jobs) {
                for (int i = 0; i < n; i++) {
                    if (idl[i] != 4) continue;
                    for (int j = 0; j < n; j++) {
                        const int v = idl[j];
                        if (v < 1 || v > 3) continue;
                        if (!((S.detected_prev[v - 1] == 1 && S.diff[v] < diff_max) || first)) continue;
                        if (nj >= job_cap) return APSE_ERR_CAPACITY;
                        apse_seq_job &J = jobs[nj];
                        memset(&J, 0, sizeof J);
                        J.frame = frame0 + fr; J.kind = v;
                        for (int a = 0; a < 3; a++) {
                            J.rvec[a] = frv[3 * j + a];
                            J.tvec[a] = ftv[3 * j + a] * tscale / S.size_corr[v];
                        }
                        for (int a = 0; a < 4; a++) J.dim[a] = dims[v][a];
                        J.src[0] = (float)S.cxy[4][0]; J.src[1] = (float)S.cxy[4][1];
                        J.tgt[0] = (float)S.cxy[v][0]; J.tgt[1] = (float)S.cxy[v][1];
                        J.scale = S.marker_length / ((S.msp[4] + S.msp[v]) / 2);
                        if (row) row->job_dist[v - 1] = nj;
                        nj++;
                    }
                }
            }
            for (int a = 0; a < 4; a++) S.detected_prev[a] = detected[a];                  // :782 (inside the `if ids` block)
        }
        if (row) {
            for (int a = 0; a < 4; a++) row->detected[a] = detected[a];
            if (detected[3] == 1 && S.have_msp[4]) {
                row->host_fields = 1;
                row->marker_length = py_round(S.marker_length, 5);
                row->altitude = py_round(S.altitude, 2);
                row->fov_width = py_round((double)c.width * S.marker_length / S.msp[4], 2);
                row->fov_height = py_round((double)c.height * S.marker_length / S.msp[4], 2);
            }
        }
    }
    if (n_jobs) *n_jobs = nj;
    return APSE_OK;
}

int apse_sequence_finish(int n_frames, apse_seq_row *rows, const apse_seq_job_result *results, int n_jobs)
{
    apse_seq_state st;
    memset(&st, 0, sizeof st);
    reinterpret_cast<SeqState *>(&st)->initialised = 1;
    return apse_sequence_finish_chunk(&st, n_frames, rows, results, n_jobs);
}

int apse_sequence_finish_chunk(apse_seq_state *state, int n_frames, apse_seq_row *rows, const apse_seq_job_result *results, int n_jobs)
{
    if (!state || n_frames < 0 || (n_frames > 0 && !rows) || (n_jobs > 0 && !results)) return APSE_ERR_INVALID_ARG;
    SeqState &S = *reinterpret_cast<SeqState *>(state);
    int &leds = S.leds;
    double (&dist)[3][2] = S.dist;
    for (int fr = 0; fr < n_frames; fr++) {
        apse_seq_row &r = rows[fr];
        if (r.job_led >= 0 && r.job_led < n_jobs) leds = results[r.job_led].leds;
        for (int v = 0; v < 3; v++)
            if (r.job_dist[v] >= 0 && r.job_dist[v] < n_jobs) {
                dist[v][0] = results[r.job_dist[v]].dist_aruco;
                dist[v][1] = results[r.job_dist[v]].dist_bbox;
            }
        r.leds = r.host_fields ? leds : 0;
        for (int v = 0; v < 3; v++) {
            r.dist_aruco[v] = r.detected[v] ? py_round(dist[v][0], 3) : 0;
            r.dist_bbox[v] = r.detected[v] ? py_round(dist[v][1], 3) : 0;
        }
    }
    return APSE_OK;
}

}  // extern "C"

// str(float) of Python: shortest round-trip digits; fixed notation for 1e-4 <= |x| < 1e16 with a trailing ".0" on integers
static char *py_float_str(char *p, char *end, double x)
{
    if (x == 0) { const char *z = signbit(x) ? "-0.0" : "0.0"; size_t n = strlen(z); if (p + n <= end) { memcpy(p, z, n); p += n; } return p; }
    const double a = fabs(x);
    if (a >= 1e-4 && a < 1e9 && end - p >= 32) {
        // the CSV columns are round(value, <= 5): x is the double nearest to a decimal q / 10^nd with few digits.  A decimal of at most
        // 15 significant digits that converts to x IS Python's shortest round-trip repr of x (two different such decimals never share
        // a double), so the smallest nd with (double)q / 10^nd == x -- one correctly rounded division of exact integers -- gives the
        // text without a general shortest-digits search.  Anything else falls through to std::to_chars.
        static const double P10[7] = {1, 10, 100, 1000, 10000, 100000, 1000000};
        for (int nd = 0; nd <= 6; nd++) {
            const double sc = a * P10[nd];
            const unsigned long long q = (unsigned long long)(sc + 0.5);
            if ((double)q / P10[nd] != a) continue;
            if (x < 0) *p++ = '-';
            char tmp[24];
            int len = 0;
            unsigned long long v = q;
            do { tmp[len++] = (char)('0' + v % 10); v /= 10; } while (v);
            while (len <= nd) tmp[len++] = '0';            // at least one digit before the point
            for (int i = len - 1; i >= nd; i--) *p++ = tmp[i];
            *p++ = '.';
            if (nd == 0) *p++ = '0';
            for (int i = nd - 1; i >= 0; i--) *p++ = tmp[i];
            return p;
        }
    }
    if (a >= 1e-4 && a < 1e16) {
        auto r = std::to_chars(p, end, x, std::chars_format::fixed);
        bool dot = false;
        for (char *q = p; q < r.ptr; q++) dot |= *q == '.';
        p = r.ptr;
        if (!dot && p + 2 <= end) { *p++ = '.'; *p++ = '0'; }
        return p;
    }
    auto r = std::to_chars(p, end, x, std::chars_format::scientific);
    return r.ptr;
}

extern "C" int64_t apse_sequence_csv(const apse_seq_row *rows, int n_frames, int with_header, char *buf, int64_t cap)
{
    if (!rows || !buf || cap <= 0 || n_frames < 0) return APSE_ERR_INVALID_ARG;
    static const char *HEADER = "frame_ID ,ID_4_detected ,markerLength ,leds_ID ,UAV_altitude ,fov_width ,fov_height ,"
                                "ID_1_detected ,distance_veh1_aruco ,distance_veh1_aruco_bbox ,"
                                "ID_2_detected ,distance_veh2_aruco ,distance_veh2_aruco_bbox ,"
                                "ID_3_detected ,distance_veh3_aruco ,distance_veh3_aruco_bbox ,\n";   // :136-139
    char *p = buf, *end = buf + cap;
    if (with_header) { size_t n = strlen(HEADER); if (p + n > end) return APSE_ERR_CAPACITY; memcpy(p, HEADER, n); p += n; }
    auto put_int = [&](long long v) { auto r = std::to_chars(p, end, v); p = r.ptr; };
    auto put_c = [&](char ch) { if (p < end) *p++ = ch; };
    for (int fr = 0; fr < n_frames; fr++) {
        if (end - p < 512) return APSE_ERR_CAPACITY;
        const apse_seq_row &r = rows[fr];
        put_int(r.frame_id); put_c(',');
        put_int(r.detected[3]); put_c(',');
        if (r.host_fields) {
            p = py_float_str(p, end, r.marker_length); put_c(',');
            put_int(r.leds); put_c(',');
            p = py_float_str(p, end, r.altitude); put_c(',');
            p = py_float_str(p, end, r.fov_width); put_c(',');
            p = py_float_str(p, end, r.fov_height);
        } else {
            memcpy(p, "0,0,0,0,0", 9); p += 9;   // the reference writes integer zeros here (:160-166)
        }
        for (int v = 0; v < 3; v++) {
            put_c(',');
            if (r.detected[v]) {
                put_c('1'); put_c(',');
                p = py_float_str(p, end, r.dist_aruco[v]); put_c(',');
                p = py_float_str(p, end, r.dist_bbox[v]);
            } else {
                memcpy(p, "0,0,0", 5); p += 5;
            }
        }
        put_c('\n');
    }
    if (p < end) *p = 0;
    return (int64_t)(p - buf);
}

// ---------------------------------------------------------------------------------------------------------
// device: one warp per job
struct SeqCam { double fx, fy, cx, cy, k[12]; };

__device__ __forceinline__ void seq_rodrigues(const double r[3], double R[9])
{
    const double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (theta < DBL_EPSILON) { for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1. : 0.; return; }
    const double c = cos(theta), s = sin(theta), c1 = 1. - c, it = 1. / theta;
    const double rx = r[0] * it, ry = r[1] * it, rz = r[2] * it;
    const double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
    const double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
    for (int k = 0; k < 9; k++) R[k] = c * ((k % 4 == 0) ? 1. : 0.) + c1 * rrt[k] + s * r_x[k];
}

// cv2.projectPoints of one point (same expression order as project_points in pose.cu)
__device__ __forceinline__ void seq_project(const double R[9], const double t[3], const SeqCam &C, double X, double Y, double Z, double &u, double &v)
{
    const double *k = C.k;
    double x = R[0] * X + R[1] * Y + R[2] * Z + t[0];
    double y = R[3] * X + R[4] * Y + R[5] * Z + t[1];
    double z = R[6] * X + R[7] * Y + R[8] * Z + t[2];
    z = z ? 1. / z : 1;
    x *= z; y *= z;
    const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    const double a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
    const double cdist = 1 + k[0] * r2 + k[1] * r4 + k[4] * r6;
    const double icdist2 = 1. / (1 + k[5] * r2 + k[6] * r4 + k[7] * r6);
    const double xd = x * cdist * icdist2 + k[2] * a1 + k[3] * a2 + k[8] * r2 + k[9] * r4;
    const double yd = y * cdist * icdist2 + k[2] * a3 + k[3] * a1 + k[10] * r2 + k[11] * r4;
    u = xd * C.fx + C.cx;
    v = yd * C.fy + C.cy;
}

// np.maximum(0, np.int32(v)): truncation toward zero, then clamp (:345,378,425,469)
__device__ __forceinline__ int seq_px(double v) { const int i = __double2int_rz(v); return i > 0 ? i : 0; }

__device__ __forceinline__ void seq_np_slice(int start, int stop, int len, int &a, int &b)
{
    if (start < 0) { start += len; if (start < 0) start = 0; }
    if (stop < 0) { stop += len; if (stop < 0) stop = 0; }
    a = min(start, len); b = min(stop, len);
}

// np.linspace(a, b, num)[i]: arange * step + start, the last sample set to stop exactly
__device__ __forceinline__ double seq_linspace(double a, double b, int num, int i)
{
    if (i == num - 1) return b;
    const double step = (b - a) / (double)(num - 1);
    return (double)i * step + a;
}

__global__ void __launch_bounds__(128) k_sequence_jobs(const apse_seq_job *__restrict__ jobs, int n_jobs, const uint8_t *__restrict__ gray,
                                                       int frame0, int n_gray_frames, int w, int h, SeqCam C,
                                                       apse_seq_job_result *__restrict__ out)
{
    const int ji = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (ji >= n_jobs) return;
    const apse_seq_job J = jobs[ji];
    double R[9];
    seq_rodrigues(J.rvec, R);
    if (J.kind == 0) {
        // LED strip (:338-373): lanes 0..7 project one LED each and sum its 5x5 neighbourhood with numpy's slicing rules
        const int lf = J.frame - frame0;
        if (!gray || lf < 0 || lf >= n_gray_frames) return;   // frame owned by another rank: its owner fills this result
        bool on = false;
        if (lane < 8) {
            double u, v;
            seq_project(R, J.tvec, C, (double)c_led_axis[lane][0], (double)c_led_axis[lane][1], 0.0, u, v);
            const int x = seq_px(u), y = seq_px(v);
            int x0, x1, y0, y1;
            seq_np_slice(x - 2, x + 3, w, x0, x1);
            seq_np_slice(y - 2, y + 3, h, y0, y1);
            const uint8_t *g = gray + (size_t)lf * w * h;
            long long s = 0;
            for (int yy = y0; yy < y1; yy++)
                for (int xx = x0; xx < x1; xx++) s += g[(size_t)yy * w + xx];
            on = (double)s / 25 > (double)J.led_threshold;
        }
        const unsigned b = __ballot_sync(0xffffffffu, on) & 0xffu;
        if (lane == 0) {
            int leds = 0;
            for (int j = 0; j < 8; j++) if (b >> j & 1) leds += 1 << (7 - j);
            out[ji].leds = leds;
            out[ji].valid = 1;
        }
        return;
    }
    // nearest point of the vehicle outline (:433-481): 20 + 20 + 8 + 8 points, lanes take two each
    double best = INFINITY;
    int best_i = 1 << 30, best_x = 0, best_y = 0;
    for (int i = lane; i < 56; i += 32) {
        double X, Y;
        if (i < 20) { X = J.dim[2]; Y = seq_linspace(J.dim[0], J.dim[1], 20, i); }
        else if (i < 40) { X = J.dim[3]; Y = seq_linspace(J.dim[0], J.dim[1], 20, i - 20); }
        else if (i < 48) { X = seq_linspace(J.dim[2], J.dim[3], 8, i - 40); Y = J.dim[0]; }
        else { X = seq_linspace(J.dim[2], J.dim[3], 8, i - 48); Y = J.dim[1]; }
        double u, v;
        seq_project(R, J.tvec, C, X, Y, 0.0, u, v);
        const int px = seq_px(u), py = seq_px(v);
        const double dx = (double)J.src[0] - (double)px, dy = (double)J.src[1] - (double)py;
        const double d = sqrt(dx * dx + dy * dy);
        if (d < best) { best = d; best_i = i; best_x = px; best_y = py; }   // strict <: the first minimum wins
        // the outline corners drawContours connects (:421-425): points 0, 19, 39, 20 of the sampled outline
        const int corner = i == 0 ? 0 : i == 19 ? 1 : i == 39 ? 2 : i == 20 ? 3 : -1;
        if (corner >= 0) { out[ji].outline_px[corner][0] = px; out[ji].outline_px[corner][1] = py; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o), ox = __shfl_xor_sync(0xffffffffu, best_x, o), oy = __shfl_xor_sync(0xffffffffu, best_y, o);
        if (ob < best || (ob == best && oi < best_i)) { best = ob; best_i = oi; best_x = ox; best_y = oy; }
    }
    if (lane == 0) {
        // calculateDistance (:483-492): marker-to-marker distance in float32 (both operands are float32 arrays), marker-to-outline in float64
        const float ax = J.src[0] - J.tgt[0], ay = J.src[1] - J.tgt[1];
        const float d_aruco = sqrtf(ax * ax + ay * ay);
        const double bx = (double)J.src[0] - (double)best_x, by = (double)J.src[1] - (double)best_y;
        const double d_bbox = sqrt(bx * bx + by * by);
        out[ji].dist_aruco = (double)d_aruco * J.scale;
        out[ji].dist_bbox = d_bbox * J.scale;
        out[ji].nearest_px[0] = best_x; out[ji].nearest_px[1] = best_y;
        out[ji].valid = 1;
    }
}

extern "C" int apse_sequence_jobs(apse_ctx *ctx, const apse_seq_job *jobs_host, int n_jobs, const uint8_t *gray, int frame0,
                                  int n_gray_frames, int w, int h, const double K[9], const double D[14],
                                  apse_seq_job_result *results_host, void *stream)
{
    if (!ctx || n_jobs < 0 || (n_jobs > 0 && (!jobs_host || !results_host)) || !K || !D) CTX_FAIL(ctx, APSE_ERR_INVALID_ARG, "sequence_jobs: bad argument");
    if (n_jobs == 0) return APSE_OK;
    if (D[12] != 0 || D[13] != 0) CTX_FAIL(ctx, APSE_ERR_UNSUPPORTED, "tilted sensor model (tauX/tauY) is not supported");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_jobs > ctx->seq_cap) {
        cudaFree(ctx->seq_jobs); cudaFree(ctx->seq_results);
        ctx->seq_jobs = nullptr; ctx->seq_results = nullptr; ctx->seq_cap = 0;
        const int cap = n_jobs + n_jobs / 2 + 64;
        CUDA_TRY(ctx, cudaMalloc(&ctx->seq_jobs, (size_t)cap * sizeof(apse_seq_job)));
        CUDA_TRY(ctx, cudaMalloc(&ctx->seq_results, (size_t)cap * sizeof(apse_seq_job_result)));
        ctx->seq_cap = cap;
    }
    SeqCam C;
    C.fx = K[0]; C.fy = K[4]; C.cx = K[2]; C.cy = K[5];
    for (int i = 0; i < 12; i++) C.k[i] = D[i];
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->seq_jobs, jobs_host, (size_t)n_jobs * sizeof(apse_seq_job), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->seq_results, 0, (size_t)n_jobs * sizeof(apse_seq_job_result), st));
    KLAUNCH(ctx, KID_SEQ_JOBS, st, k_sequence_jobs<<<div_up(n_jobs * 32, 128), 128, 0, st>>>((const apse_seq_job *)ctx->seq_jobs, n_jobs, gray, frame0, n_gray_frames, w, h, C,
                                                                                            (apse_seq_job_result *)ctx->seq_results));
    CUDA_TRY(ctx, cudaMemcpyAsync(results_host, ctx->seq_results, (size_t)n_jobs * sizeof(apse_seq_job_result), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return APSE_OK;
}
