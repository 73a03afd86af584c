/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the pose stage.
 *
 *   aruco_detect.py:601          aruco.estimatePoseSingleMarkers(corners, markerLength, mtx, dist)
 *                                = per marker solvePnP(ITERATIVE) on the 4 float32 object points
 *   aruco_detect.py:344,377,424,468   cv2.projectPoints(obj, rvec, tvec/size_corr, mtx, dist)
 *
 * The arithmetic lives in OpenCV calib3d (un-vendored dependency, reference README.md:42).  Restated from
 * the published algorithm (Zhang planar initialisation + Levenberg-Marquardt on the reprojection error,
 * rational distortion model) as specified in SURVEY.md Appendix A.8; pinned against cv2 4.13.0
 * (cv2.solvePnP / cv2.projectPoints / cv2.undistortPoints / cv2.Rodrigues) by tests/test_oracle_pose.py.
 */
#include <math.h>
#include <float.h>
#include <string.h>
#include <stdint.h>

/* ---------------------------------------------------------------------------------------------------- */
static void rodrigues_vec2mat(const double r_in[3], double R[9], double J[27] /* nullable, 3x9 */)
{
    double rx = r_in[0], ry = r_in[1], rz = r_in[2];
    double theta = sqrt(rx * rx + ry * ry + rz * rz);
    if (theta < DBL_EPSILON) {
        memset(R, 0, 9 * sizeof(double));
        R[0] = R[4] = R[8] = 1;
        if (J) {
            memset(J, 0, 27 * sizeof(double));
            J[5] = J[15] = J[19] = -1;
            J[7] = J[11] = J[21] = 1;
        }
        return;
    }
    double c = cos(theta), s = sin(theta), c1 = 1. - c, itheta = 1. / theta;
    rx *= itheta; ry *= itheta; rz *= itheta;
    double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
    double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
    static const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int k = 0; k < 9; k++) R[k] = c * I[k] + c1 * rrt[k] + s * r_x[k];
    if (J) {
        double drrt[27] = {rx + rx, ry, rz, ry, 0, 0, rz, 0, 0, 0, rx, 0, rx, ry + ry, rz, 0, rz, 0,
                           0, 0, rx, 0, 0, ry, rx, ry, rz + rz};
        static const double d_r_x[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 1, 0, 0, 0, -1, 0, 0,
                                         0, -1, 0, 1, 0, 0, 0, 0, 0};
        double rv[3] = {rx, ry, rz};
        for (int i = 0; i < 3; i++) {
            double ri = rv[i];
            double a0 = -s * ri, a1 = (s - 2 * c1 * itheta) * ri, a2 = c1 * itheta;
            double a3 = (c - s * itheta) * ri, a4 = s * itheta;
            for (int k = 0; k < 9; k++)
                J[i * 9 + k] = a0 * I[k] + a1 * rrt[k] + a2 * drrt[i * 9 + k] + a3 * r_x[k] + a4 * d_r_x[i * 9 + k];
        }
    }
}

static void inv3(const double *m, double *o)
{
    double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    d = 1. / d;
    o[0] = (m[4] * m[8] - m[5] * m[7]) * d; o[1] = (m[2] * m[7] - m[1] * m[8]) * d; o[2] = (m[1] * m[5] - m[2] * m[4]) * d;
    o[3] = (m[5] * m[6] - m[3] * m[8]) * d; o[4] = (m[0] * m[8] - m[2] * m[6]) * d; o[5] = (m[2] * m[3] - m[0] * m[5]) * d;
    o[6] = (m[3] * m[7] - m[4] * m[6]) * d; o[7] = (m[1] * m[6] - m[0] * m[7]) * d; o[8] = (m[0] * m[4] - m[1] * m[3]) * d;
}

/* orthogonal polar factor U*Vt of a near-rotation matrix (what the dependency gets from its SVD) */
static void orthonormalize(double R[9])
{
    for (int it = 0; it < 30; it++) {
        double Ri[9], N[9], diff = 0;
        inv3(R, Ri);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                N[i * 3 + j] = 0.5 * (R[i * 3 + j] + Ri[j * 3 + i]);
                diff += fabs(N[i * 3 + j] - R[i * 3 + j]);
            }
        memcpy(R, N, sizeof N);
        if (diff < 1e-15) break;
    }
}

static void rodrigues_mat2vec(const double Rin[9], double r[3])
{
    double R[9];
    memcpy(R, Rin, sizeof R);
    orthonormalize(R);
    double x = R[7] - R[5], y = R[2] - R[6], z = R[3] - R[1];
    double s = sqrt((x * x + y * y + z * z) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1. ? 1. : c < -1. ? -1. : c;
    double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) { r[0] = r[1] = r[2] = 0; return; }
        double t;
        t = (R[0] + 1) * 0.5; x = sqrt(t > 0 ? t : 0);
        t = (R[4] + 1) * 0.5; y = sqrt(t > 0 ? t : 0) * (R[1] < 0 ? -1. : 1.);
        t = (R[8] + 1) * 0.5; z = sqrt(t > 0 ? t : 0) * (R[2] < 0 ? -1. : 1.);
        if (fabs(x) < fabs(y) && fabs(x) < fabs(z) && (R[5] > 0) != (y * z > 0)) z = -z;
        theta /= sqrt(x * x + y * y + z * z);
        r[0] = x * theta; r[1] = y * theta; r[2] = z * theta;
    } else {
        double vth = 1 / (2 * s);
        vth *= theta;
        r[0] = x * vth; r[1] = y * vth; r[2] = z * vth;
    }
}

void orc_rodrigues(const double *r, double *R, double *J) { rodrigues_vec2mat(r, R, J); }
void orc_rodrigues_inv(const double *R, double *r) { rodrigues_mat2vec(R, r); }

/* ---------------------------------------------------------------------------------------------------- */
/* projectPoints, 14-coefficient model (tilt terms must be 0).  dpdr/dpdt: [2n][3] row-major, nullable */
void orc_project_points(const double *obj, int n, const double *rvec, const double *tvec, const double *K,
                        const double *k, double *img, double *dpdr, double *dpdt)
{
    double R[9], dRdr[27];
    rodrigues_vec2mat(rvec, R, dpdr ? dRdr : NULL);
    double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    for (int i = 0; i < n; i++) {
        double X = obj[3 * i], Y = obj[3 * i + 1], Z = obj[3 * i + 2];
        double x = R[0] * X + R[1] * Y + R[2] * Z + tvec[0];
        double y = R[3] * X + R[4] * Y + R[5] * Z + tvec[1];
        double z = R[6] * X + R[7] * Y + R[8] * Z + tvec[2];
        z = z ? 1. / z : 1;
        x *= z; y *= z;
        double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
        double a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
        double cdist = 1 + k[0] * r2 + k[1] * r4 + k[4] * r6;
        double icdist2 = 1. / (1 + k[5] * r2 + k[6] * r4 + k[7] * r6);
        double xd = x * cdist * icdist2 + k[2] * a1 + k[3] * a2 + k[8] * r2 + k[9] * r4;
        double yd = y * cdist * icdist2 + k[2] * a3 + k[3] * a1 + k[10] * r2 + k[11] * r4;
        img[2 * i] = xd * fx + cx;
        img[2 * i + 1] = yd * fy + cy;
        if (dpdt) {
            double dxdt[3] = {z, 0, -x * z}, dydt[3] = {0, z, -y * z};
            for (int j = 0; j < 3; j++) {
                double dr2dt = 2 * x * dxdt[j] + 2 * y * dydt[j];
                double dcdist_dt = k[0] * dr2dt + 2 * k[1] * r2 * dr2dt + 3 * k[4] * r4 * dr2dt;
                double dicdist2_dt = -icdist2 * icdist2 * (k[5] * dr2dt + 2 * k[6] * r2 * dr2dt + 3 * k[7] * r4 * dr2dt);
                double da1dt = 2 * (x * dydt[j] + y * dxdt[j]);
                double dmxdt = dxdt[j] * cdist * icdist2 + x * dcdist_dt * icdist2 + x * cdist * dicdist2_dt +
                               k[2] * da1dt + k[3] * (dr2dt + 4 * x * dxdt[j]) + k[8] * dr2dt + 2 * r2 * k[9] * dr2dt;
                double dmydt = dydt[j] * cdist * icdist2 + y * dcdist_dt * icdist2 + y * cdist * dicdist2_dt +
                               k[2] * (dr2dt + 4 * y * dydt[j]) + k[3] * da1dt + k[10] * dr2dt + 2 * r2 * k[11] * dr2dt;
                dpdt[(2 * i) * 3 + j] = fx * dmxdt;
                dpdt[(2 * i + 1) * 3 + j] = fy * dmydt;
            }
        }
        if (dpdr) {
            double dx0dr[3], dy0dr[3], dz0dr[3];
            for (int j = 0; j < 3; j++) {
                dx0dr[j] = X * dRdr[j * 9 + 0] + Y * dRdr[j * 9 + 1] + Z * dRdr[j * 9 + 2];
                dy0dr[j] = X * dRdr[j * 9 + 3] + Y * dRdr[j * 9 + 4] + Z * dRdr[j * 9 + 5];
                dz0dr[j] = X * dRdr[j * 9 + 6] + Y * dRdr[j * 9 + 7] + Z * dRdr[j * 9 + 8];
            }
            for (int j = 0; j < 3; j++) {
                double dxdr = z * (dx0dr[j] - x * dz0dr[j]);
                double dydr = z * (dy0dr[j] - y * dz0dr[j]);
                double dr2dr = 2 * x * dxdr + 2 * y * dydr;
                double dcdist_dr = (k[0] + 2 * k[1] * r2 + 3 * k[4] * r4) * dr2dr;
                double dicdist2_dr = -icdist2 * icdist2 * (k[5] + 2 * k[6] * r2 + 3 * k[7] * r4) * dr2dr;
                double da1dr = 2 * (x * dydr + y * dxdr);
                double dmxdr = dxdr * cdist * icdist2 + x * dcdist_dr * icdist2 + x * cdist * dicdist2_dr +
                               k[2] * da1dr + k[3] * (dr2dr + 4 * x * dxdr) + (k[8] + 2 * r2 * k[9]) * dr2dr;
                double dmydr = dydr * cdist * icdist2 + y * dcdist_dr * icdist2 + y * cdist * dicdist2_dr +
                               k[2] * (dr2dr + 4 * y * dydr) + k[3] * da1dr + (k[10] + 2 * r2 * k[11]) * dr2dr;
                dpdr[(2 * i) * 3 + j] = fx * dmxdr;
                dpdr[(2 * i + 1) * 3 + j] = fy * dmydr;
            }
        }
    }
}

/* undistortPoints to normalised coordinates, exactly 5 fixed-point iterations */
void orc_undistort_points(const double *pts, int n, const double *K, const double *k, double *out)
{
    double fx = K[0], fy = K[4], cx = K[2], cy = K[5], ifx = 1. / fx, ify = 1. / fy;
    for (int i = 0; i < n; i++) {
        double u = pts[2 * i], v = pts[2 * i + 1];
        double x = (u - cx) * ifx, y = (v - cy) * ify, x0 = x, y0 = y;
        for (int j = 0; j < 5; j++) {
            double r2 = x * x + y * y;
            double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
            double dX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
            double dY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
            x = (x0 - dX) * icdist;
            y = (y0 - dY) * icdist;
        }
        out[2 * i] = x;
        out[2 * i + 1] = y;
    }
}

/* ---------------------------------------------------------------------------------------------------- */
static int lu_solve(double *A, double *b, int n)
{
    for (int i = 0; i < n; i++) {
        int k = i;
        for (int j = i + 1; j < n; j++) if (fabs(A[j * n + i]) > fabs(A[k * n + i])) k = j;
        if (fabs(A[k * n + i]) < 1e-300) return 0;
        if (k != i) {
            for (int j = 0; j < n; j++) { double t = A[i * n + j]; A[i * n + j] = A[k * n + j]; A[k * n + j] = t; }
            double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        for (int j = i + 1; j < n; j++) {
            double f = A[j * n + i] / A[i * n + i];
            for (int c = i; c < n; c++) A[j * n + c] -= f * A[i * n + c];
            b[j] -= f * b[i];
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = b[i];
        for (int c = i + 1; c < n; c++) s -= A[i * n + c] * b[c];
        b[i] = s / A[i * n + i];
    }
    return 1;
}

/* exact 4-point homography src(x,y) -> dst(x,y), h33 = 1 */
static int homography4(const double *src, const double *dst, double *H)
{
    double A[64], b[8];
    memset(A, 0, sizeof A);
    for (int i = 0; i < 4; i++) {
        double X = src[2 * i], Y = src[2 * i + 1], x = dst[2 * i], y = dst[2 * i + 1];
        double *r0 = A + i * 8, *r1 = A + (i + 4) * 8;
        r0[0] = X; r0[1] = Y; r0[2] = 1; r0[6] = -x * X; r0[7] = -x * Y; b[i] = x;
        r1[3] = X; r1[4] = Y; r1[5] = 1; r1[6] = -y * X; r1[7] = -y * Y; b[i + 4] = y;
    }
    if (!lu_solve(A, b, 8)) return 0;
    memcpy(H, b, 8 * sizeof(double));
    H[8] = 1;
    return 1;
}

/* symmetric eigen-decomposition (cyclic Jacobi): A = V diag(w) V^T, V columns */
static void jacobi_eig(double *A, int n, double *w, double *V)
{
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i * n + j] = i == j;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j];
        if (off < 1e-300) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                if (fabs(A[p * n + q]) < 1e-300) continue;
                double th = (A[q * n + q] - A[p * n + p]) / (2 * A[p * n + q]);
                double t = (th >= 0 ? 1. : -1.) / (fabs(th) + sqrt(th * th + 1));
                double c = 1 / sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < n; k++) {
                    double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq;
                    V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < n; i++) w[i] = A[i * n + i];
}

/* x = pinv(A) b for symmetric A (the dependency solves the damped normal equations by SVD) */
static void sym_solve_svd(const double *A_in, const double *b, int n, double *x)
{
    double A[36], V[36], w[6], thr = 0;
    memcpy(A, A_in, sizeof(double) * n * n);
    jacobi_eig(A, n, w, V);
    for (int i = 0; i < n; i++) thr += fabs(w[i]);
    thr *= 2 * DBL_EPSILON;
    for (int i = 0; i < n; i++) x[i] = 0;
    for (int k = 0; k < n; k++) {
        if (fabs(w[k]) <= thr) continue;
        double s = 0;
        for (int i = 0; i < n; i++) s += V[i * n + k] * b[i];
        s /= w[k];
        for (int i = 0; i < n; i++) x[i] += V[i * n + k] * s;
    }
}

static double norm_n(const double *a, int n)
{
    double s = 0;
    for (int i = 0; i < n; i++) s += a[i] * a[i];
    return sqrt(s);
}

/* solvePnP(ITERATIVE) for 4 coplanar (z = 0, centred) object points.
 * obj: 4x3 doubles (already float32-rounded by the caller), img: 4x2 doubles. Returns LM iteration count. */
int orc_solve_pnp_planar(const double *obj, const double *img, const double *K, const double *k, double *rvec,
                         double *tvec)
{
    const int n = 4;
    double mn[8], Mxy[8], H[9];
    orc_undistort_points(img, n, K, k, mn);
    for (int i = 0; i < n; i++) { Mxy[2 * i] = obj[3 * i]; Mxy[2 * i + 1] = obj[3 * i + 1]; }
    double param[6] = {0, 0, 0, 0, 0, 0};
    if (homography4(Mxy, mn, H)) {
        double h1n = sqrt(H[0] * H[0] + H[3] * H[3] + H[6] * H[6]);
        double h2n = sqrt(H[1] * H[1] + H[4] * H[4] + H[7] * H[7]);
        double s1 = 1. / fmax(h1n, DBL_EPSILON), s2 = 1. / fmax(h2n, DBL_EPSILON);
        double st = 2. / fmax(h1n + h2n, DBL_EPSILON);
        double h1[3] = {H[0] * s1, H[3] * s1, H[6] * s1}, h2[3] = {H[1] * s2, H[4] * s2, H[7] * s2};
        double t[3] = {H[2] * st, H[5] * st, H[8] * st};
        double h3[3] = {h1[1] * h2[2] - h1[2] * h2[1], h1[2] * h2[0] - h1[0] * h2[2], h1[0] * h2[1] - h1[1] * h2[0]};
        double R0[9] = {h1[0], h2[0], h3[0], h1[1], h2[1], h3[1], h1[2], h2[2], h3[2]};
        double r[3], R[9];
        rodrigues_mat2vec(R0, r);
        rodrigues_vec2mat(r, R, NULL);
        rodrigues_mat2vec(R, r);
        memcpy(param, r, sizeof r);
        memcpy(param + 3, t, sizeof t);
    }
    /* Levenberg-Marquardt, multiplicative damping, <= 20 accepted iterations, eps = FLT_EPSILON */
    double prev[6], J[8 * 6], err[8], JtJ[36], JtErr[6], proj[8], dpdr[24], dpdt[24];
    double prevErrNorm = DBL_MAX, errNorm;
    int lambdaLg10 = -3, iters = 0;
    const double LOG10 = log(10.);
    for (;;) {
        /* CALC_J at param */
        orc_project_points(obj, n, param, param + 3, K, k, proj, dpdr, dpdt);
        for (int i = 0; i < 2 * n; i++) {
            err[i] = proj[i] - img[i];
            for (int j = 0; j < 3; j++) { J[i * 6 + j] = dpdr[i * 3 + j]; J[i * 6 + 3 + j] = dpdt[i * 3 + j]; }
        }
        for (int a = 0; a < 6; a++) {
            for (int b = 0; b < 6; b++) {
                double s = 0;
                for (int i = 0; i < 2 * n; i++) s += J[i * 6 + a] * J[i * 6 + b];
                JtJ[a * 6 + b] = s;
            }
            double s = 0;
            for (int i = 0; i < 2 * n; i++) s += J[i * 6 + a] * err[i];
            JtErr[a] = s;
        }
        memcpy(prev, param, sizeof prev);
        if (iters == 0) prevErrNorm = norm_n(err, 2 * n);
        for (;;) {
            /* step */
            double A[36], d[6], lambda = exp(lambdaLg10 * LOG10);
            memcpy(A, JtJ, sizeof A);
            for (int a = 0; a < 6; a++) A[a * 6 + a] *= 1. + lambda;
            sym_solve_svd(A, JtErr, 6, d);
            for (int a = 0; a < 6; a++) param[a] = prev[a] - d[a];
            /* CHECK_ERR */
            orc_project_points(obj, n, param, param + 3, K, k, proj, NULL, NULL);
            for (int i = 0; i < 2 * n; i++) err[i] = proj[i] - img[i];
            errNorm = norm_n(err, 2 * n);
            if (errNorm > prevErrNorm) {
                if (++lambdaLg10 <= 16) continue;
            }
            break;
        }
        lambdaLg10 = lambdaLg10 - 1 > -16 ? lambdaLg10 - 1 : -16;
        double dd[6];
        for (int a = 0; a < 6; a++) dd[a] = param[a] - prev[a];
        if (++iters >= 20 || norm_n(dd, 6) / norm_n(prev, 6) < FLT_EPSILON) break;
        prevErrNorm = errNorm;
    }
    memcpy(rvec, param, 3 * sizeof(double));
    memcpy(tvec, param + 3, 3 * sizeof(double));
    return iters;
}

/* estimatePoseSingleMarkers: corners [n][4][2] float32, marker_length cast to float as the legacy API did */
void orc_estimate_pose_single_markers(const float *corners, int n, float marker_length, const double *K,
                                      const double *k, double *rvecs, double *tvecs)
{
    float h = marker_length / 2.f;
    double obj[12] = {-h, h, 0, h, h, 0, h, -h, 0, -h, -h, 0};
    for (int i = 0; i < n; i++) {
        double img[8];
        for (int j = 0; j < 8; j++) img[j] = corners[8 * i + j];
        orc_solve_pnp_planar(obj, img, K, k, rvecs + 3 * i, tvecs + 3 * i);
    }
}
