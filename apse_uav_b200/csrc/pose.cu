// pose.cu -- batched per-marker pose (FP64, four lanes per marker) and point projection (one thread per point):
//   k_pose4 / k_pose_frames4   aruco_detect.py:601  aruco.estimatePoseSingleMarkers = solvePnP(ITERATIVE) per marker
//   k_project_points  aruco_detect.py:344,377,424,468  cv2.projectPoints with the 14-coefficient model
// Recipe of the dependency: 5 fixed-point undistortion iterations -> planar homography
// initialisation -> Levenberg-Marquardt (lambda = 10^k, diag*(1+lambda), <= 20 accepted iterations,
// relative step < FLT_EPSILON), normal equations solved through the eigen-decomposition of J^T J.
#include "common.cuh"
#include <math.h>
#include <float.h>

struct CamModel { double fx, fy, cx, cy, k[12]; };

__device__ void rodrigues_vec2mat(const double r_in[3], double R[9], double *J /* 27 or null */)
{
    double rx = r_in[0], ry = r_in[1], rz = r_in[2];
    double theta = sqrt(rx * rx + ry * ry + rz * rz);
    if (theta < DBL_EPSILON) {
        for (int i = 0; i < 9; i++) R[i] = 0;
        R[0] = R[4] = R[8] = 1;
        if (J) {
            for (int i = 0; i < 27; i++) J[i] = 0;
            J[5] = J[15] = J[19] = -1;
            J[7] = J[11] = J[21] = 1;
        }
        return;
    }
    double c = cos(theta), s = sin(theta), c1 = 1. - c, itheta = 1. / theta;
    rx *= itheta; ry *= itheta; rz *= itheta;
    const double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
    const double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int k = 0; k < 9; k++) R[k] = c * I[k] + c1 * rrt[k] + s * r_x[k];
    if (J) {
        const double drrt[27] = {rx + rx, ry, rz, ry, 0, 0, rz, 0, 0, 0, rx, 0, rx, ry + ry, rz, 0, rz, 0,
                                 0, 0, rx, 0, 0, ry, rx, ry, rz + rz};
        const double d_r_x[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 1, 0, 0, 0, -1, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0, 0};
        const double rv[3] = {rx, ry, rz};
        for (int i = 0; i < 3; i++) {
            double ri = rv[i];
            double a0 = -s * ri, a1 = (s - 2 * c1 * itheta) * ri, a2 = c1 * itheta;
            double a3 = (c - s * itheta) * ri, a4 = s * itheta;
            for (int k = 0; k < 9; k++)
                J[i * 9 + k] = a0 * I[k] + a1 * rrt[k] + a2 * drrt[i * 9 + k] + a3 * r_x[k] + a4 * d_r_x[i * 9 + k];
        }
    }
}

__device__ void inv3(const double *m, double *o)
{
    double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    d = 1. / d;
    o[0] = (m[4] * m[8] - m[5] * m[7]) * d; o[1] = (m[2] * m[7] - m[1] * m[8]) * d; o[2] = (m[1] * m[5] - m[2] * m[4]) * d;
    o[3] = (m[5] * m[6] - m[3] * m[8]) * d; o[4] = (m[0] * m[8] - m[2] * m[6]) * d; o[5] = (m[2] * m[3] - m[0] * m[5]) * d;
    o[6] = (m[3] * m[7] - m[4] * m[6]) * d; o[7] = (m[1] * m[6] - m[0] * m[7]) * d; o[8] = (m[0] * m[4] - m[1] * m[3]) * d;
}

// rotation matrix -> rotation vector; the input is first replaced by its orthogonal polar factor (U V^T)
__device__ void rodrigues_mat2vec(const double Rin[9], double r[3])
{
    double R[9];
    for (int i = 0; i < 9; i++) R[i] = Rin[i];
    for (int it = 0; it < 30; it++) {
        double Ri[9], N[9], diff = 0;
        inv3(R, Ri);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                N[i * 3 + j] = 0.5 * (R[i * 3 + j] + Ri[j * 3 + i]);
                diff += fabs(N[i * 3 + j] - R[i * 3 + j]);
            }
        for (int i = 0; i < 9; i++) R[i] = N[i];
        if (diff < 1e-15) break;
    }
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

// projects n object points; dpdr / dpdt: [2n][3] (nullable)
__device__ void project_points(const double *obj, int n, const double *rvec, const double *tvec, const CamModel &C,
                               double *img, double *dpdr, double *dpdt)
{
    double R[9], dRdr[27];
    rodrigues_vec2mat(rvec, R, dpdr ? dRdr : nullptr);
    const double *k = C.k;
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
        img[2 * i] = xd * C.fx + C.cx;
        img[2 * i + 1] = yd * C.fy + C.cy;
        if (dpdt) {
            const double dxdt[3] = {z, 0, -x * z}, dydt[3] = {0, z, -y * z};
            for (int j = 0; j < 3; j++) {
                double dr2dt = 2 * x * dxdt[j] + 2 * y * dydt[j];
                double dcdist_dt = k[0] * dr2dt + 2 * k[1] * r2 * dr2dt + 3 * k[4] * r4 * dr2dt;
                double dicdist2_dt = -icdist2 * icdist2 * (k[5] * dr2dt + 2 * k[6] * r2 * dr2dt + 3 * k[7] * r4 * dr2dt);
                double da1dt = 2 * (x * dydt[j] + y * dxdt[j]);
                double dmxdt = dxdt[j] * cdist * icdist2 + x * dcdist_dt * icdist2 + x * cdist * dicdist2_dt + k[2] * da1dt +
                               k[3] * (dr2dt + 4 * x * dxdt[j]) + k[8] * dr2dt + 2 * r2 * k[9] * dr2dt;
                double dmydt = dydt[j] * cdist * icdist2 + y * dcdist_dt * icdist2 + y * cdist * dicdist2_dt +
                               k[2] * (dr2dt + 4 * y * dydt[j]) + k[3] * da1dt + k[10] * dr2dt + 2 * r2 * k[11] * dr2dt;
                dpdt[(2 * i) * 3 + j] = C.fx * dmxdt;
                dpdt[(2 * i + 1) * 3 + j] = C.fy * dmydt;
            }
        }
        if (dpdr) {
            for (int j = 0; j < 3; j++) {
                double dx0 = X * dRdr[j * 9 + 0] + Y * dRdr[j * 9 + 1] + Z * dRdr[j * 9 + 2];
                double dy0 = X * dRdr[j * 9 + 3] + Y * dRdr[j * 9 + 4] + Z * dRdr[j * 9 + 5];
                double dz0 = X * dRdr[j * 9 + 6] + Y * dRdr[j * 9 + 7] + Z * dRdr[j * 9 + 8];
                double dxdr = z * (dx0 - x * dz0), dydr = z * (dy0 - y * dz0);
                double dr2dr = 2 * x * dxdr + 2 * y * dydr;
                double dcdist_dr = (k[0] + 2 * k[1] * r2 + 3 * k[4] * r4) * dr2dr;
                double dicdist2_dr = -icdist2 * icdist2 * (k[5] + 2 * k[6] * r2 + 3 * k[7] * r4) * dr2dr;
                double da1dr = 2 * (x * dydr + y * dxdr);
                double dmxdr = dxdr * cdist * icdist2 + x * dcdist_dr * icdist2 + x * cdist * dicdist2_dr + k[2] * da1dr +
                               k[3] * (dr2dr + 4 * x * dxdr) + (k[8] + 2 * r2 * k[9]) * dr2dr;
                double dmydr = dydr * cdist * icdist2 + y * dcdist_dr * icdist2 + y * cdist * dicdist2_dr +
                               k[2] * (dr2dr + 4 * y * dydr) + k[3] * da1dr + (k[10] + 2 * r2 * k[11]) * dr2dr;
                dpdr[(2 * i) * 3 + j] = C.fx * dmxdr;
                dpdr[(2 * i + 1) * 3 + j] = C.fy * dmydr;
            }
        }
    }
}

__device__ bool lu_solve8(double *A, double *b)
{
    const int n = 8;
    for (int i = 0; i < n; i++) {
        int k = i;
        for (int j = i + 1; j < n; j++) if (fabs(A[j * n + i]) > fabs(A[k * n + i])) k = j;
        if (fabs(A[k * n + i]) < 1e-300) return false;
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
    return true;
}

// x = pinv(A) b, symmetric 6x6, cyclic Jacobi eigen-decomposition (the dependency solves through an SVD)
__device__ void sym_solve6(const double *Ain, const double *b, double *x)
{
    const int n = 6;
    {
        // Fast path: the damped normal matrix is symmetric positive definite in every well-posed solve, where
        // a Cholesky solve equals the pseudo-inverse solution to rounding (1e-16 * cond).  Rank-deficient
        // systems fall through to the eigen-decomposition below, which reproduces the SVD's truncation.
        double Lm[36], tr = 0;
        bool ok = true;
        for (int i = 0; i < n; i++) tr += Ain[i * n + i];
        for (int j = 0; j < n && ok; j++) {
            double d = Ain[j * n + j];
            for (int k = 0; k < j; k++) d -= Lm[j * n + k] * Lm[j * n + k];
            if (!(d > tr * 1e-13)) { ok = false; break; }
            d = sqrt(d);
            Lm[j * n + j] = d;
            for (int i = j + 1; i < n; i++) {
                double s = Ain[i * n + j];
                for (int k = 0; k < j; k++) s -= Lm[i * n + k] * Lm[j * n + k];
                Lm[i * n + j] = s / d;
            }
        }
        if (ok) {
            double y[6];
            for (int i = 0; i < n; i++) {
                double s = b[i];
                for (int k = 0; k < i; k++) s -= Lm[i * n + k] * y[k];
                y[i] = s / Lm[i * n + i];
            }
            for (int i = n - 1; i >= 0; i--) {
                double s = y[i];
                for (int k = i + 1; k < n; k++) s -= Lm[k * n + i] * x[k];
                x[i] = s / Lm[i * n + i];
            }
            return;
        }
    }
    double A[36], V[36];
    for (int i = 0; i < 36; i++) { A[i] = Ain[i]; V[i] = 0; }
    for (int i = 0; i < n; i++) V[i * n + i] = 1;
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
    double thr = 0;
    for (int i = 0; i < n; i++) thr += fabs(A[i * n + i]);
    thr *= 2 * DBL_EPSILON;
    for (int i = 0; i < n; i++) x[i] = 0;
    for (int k = 0; k < n; k++) {
        double wk = A[k * n + k];
        if (fabs(wk) <= thr) continue;
        double s = 0;
        for (int i = 0; i < n; i++) s += V[i * n + k] * b[i];
        s /= wk;
        for (int i = 0; i < n; i++) x[i] += V[i * n + k] * s;
    }
}


// ---------------------------------------------------------------------------------------------------------
// Four lanes per marker (lane j of a group owns corner j).  A thread-per-marker solver keeps its arrays in local
// memory and is one long dependent FP64 chain (~0.4 ms per launch whatever the marker count); here the projection with
// its Jacobian rows, the residuals and the partial normal equations are computed per corner in registers, the 27
// normal-equation sums are combined with two butterfly shuffles (every lane of the group ends up with bit-identical
// values, so the group's control flow never diverges), and the 6x6 Cholesky solve is unrolled in registers.
__device__ __forceinline__ double gsum4(unsigned gmask, double v)
{
    v += __shfl_xor_sync(gmask, v, 1);
    v += __shfl_xor_sync(gmask, v, 2);
    return v;
}

// one object point: image point and (optionally) its 2x3 derivative blocks w.r.t. rvec (through dRdr) and tvec
template <bool JAC>
__device__ __forceinline__ void project_one(double X, double Y, double Z, const double *R, const double *dRdr, const double *tvec,
                                            const CamModel &C, double &u, double &v, double *dr, double *dt)
{
    const double *k = C.k;
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
    u = xd * C.fx + C.cx;
    v = yd * C.fy + C.cy;
    if (JAC) {
        const double dxdt[3] = {z, 0, -x * z}, dydt[3] = {0, z, -y * z};
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double dr2dt = 2 * x * dxdt[j] + 2 * y * dydt[j];
            double dcdist_dt = k[0] * dr2dt + 2 * k[1] * r2 * dr2dt + 3 * k[4] * r4 * dr2dt;
            double dicdist2_dt = -icdist2 * icdist2 * (k[5] * dr2dt + 2 * k[6] * r2 * dr2dt + 3 * k[7] * r4 * dr2dt);
            double da1dt = 2 * (x * dydt[j] + y * dxdt[j]);
            double dmxdt = dxdt[j] * cdist * icdist2 + x * dcdist_dt * icdist2 + x * cdist * dicdist2_dt + k[2] * da1dt +
                           k[3] * (dr2dt + 4 * x * dxdt[j]) + k[8] * dr2dt + 2 * r2 * k[9] * dr2dt;
            double dmydt = dydt[j] * cdist * icdist2 + y * dcdist_dt * icdist2 + y * cdist * dicdist2_dt +
                           k[2] * (dr2dt + 4 * y * dydt[j]) + k[3] * da1dt + k[10] * dr2dt + 2 * r2 * k[11] * dr2dt;
            dt[j] = C.fx * dmxdt;
            dt[3 + j] = C.fy * dmydt;
        }
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double dx0 = X * dRdr[j * 9 + 0] + Y * dRdr[j * 9 + 1] + Z * dRdr[j * 9 + 2];
            double dy0 = X * dRdr[j * 9 + 3] + Y * dRdr[j * 9 + 4] + Z * dRdr[j * 9 + 5];
            double dz0 = X * dRdr[j * 9 + 6] + Y * dRdr[j * 9 + 7] + Z * dRdr[j * 9 + 8];
            double dxdr = z * (dx0 - x * dz0), dydr = z * (dy0 - y * dz0);
            double dr2dr = 2 * x * dxdr + 2 * y * dydr;
            double dcdist_dr = (k[0] + 2 * k[1] * r2 + 3 * k[4] * r4) * dr2dr;
            double dicdist2_dr = -icdist2 * icdist2 * (k[5] + 2 * k[6] * r2 + 3 * k[7] * r4) * dr2dr;
            double da1dr = 2 * (x * dydr + y * dxdr);
            double dmxdr = dxdr * cdist * icdist2 + x * dcdist_dr * icdist2 + x * cdist * dicdist2_dr + k[2] * da1dr +
                           k[3] * (dr2dr + 4 * x * dxdr) + (k[8] + 2 * r2 * k[9]) * dr2dr;
            double dmydr = dydr * cdist * icdist2 + y * dcdist_dr * icdist2 + y * cdist * dicdist2_dr +
                           k[2] * (dr2dr + 4 * y * dydr) + k[3] * da1dr + (k[10] + 2 * r2 * k[11]) * dr2dr;
            dr[j] = C.fx * dmxdr;
            dr[3 + j] = C.fy * dmydr;
        }
    }
}

// rodrigues_vec2mat with every loop unrolled (registers only)
template <bool JAC>
__device__ __forceinline__ void rodrigues_reg(const double *r_in, double *R, double *J)
{
    double rx = r_in[0], ry = r_in[1], rz = r_in[2];
    double theta = sqrt(rx * rx + ry * ry + rz * rz);
    if (theta < DBL_EPSILON) {
#pragma unroll
        for (int i = 0; i < 9; i++) R[i] = (i == 0 || i == 4 || i == 8) ? 1. : 0.;
        if (JAC) {
#pragma unroll
            for (int i = 0; i < 27; i++) J[i] = (i == 5 || i == 15 || i == 19) ? -1. : (i == 7 || i == 11 || i == 21) ? 1. : 0.;
        }
        return;
    }
    double c = cos(theta), s = sin(theta), c1 = 1. - c, itheta = 1. / theta;
    rx *= itheta; ry *= itheta; rz *= itheta;
    const double rrt[9] = {rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz};
    const double r_x[9] = {0, -rz, ry, rz, 0, -rx, -ry, rx, 0};
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = c * I[k] + c1 * rrt[k] + s * r_x[k];
    if (JAC) {
        const double drrt[27] = {rx + rx, ry, rz, ry, 0, 0, rz, 0, 0, 0, rx, 0, rx, ry + ry, rz, 0, rz, 0,
                                 0, 0, rx, 0, 0, ry, rx, ry, rz + rz};
        const double d_r_x[27] = {0, 0, 0, 0, 0, -1, 0, 1, 0, 0, 0, 1, 0, 0, 0, -1, 0, 0, 0, -1, 0, 1, 0, 0, 0, 0, 0};
        const double rv[3] = {rx, ry, rz};
#pragma unroll
        for (int i = 0; i < 3; i++) {
            double ri = rv[i];
            double a0 = -s * ri, a1 = (s - 2 * c1 * itheta) * ri, a2 = c1 * itheta;
            double a3 = (c - s * itheta) * ri, a4 = s * itheta;
#pragma unroll
            for (int k = 0; k < 9; k++)
                J[i * 9 + k] = a0 * I[k] + a1 * rrt[k] + a2 * drrt[i * 9 + k] + a3 * r_x[k] + a4 * d_r_x[i * 9 + k];
        }
    }
}

// Cholesky solve of the damped 6x6 normal matrix held as its lower triangle in registers; false = not positive
// definite (the caller falls back to the eigen-decomposition of sym_solve6)
__device__ __forceinline__ bool chol_solve6_reg(const double *Al /* 21: row-major lower triangle */, const double *b, double *x)
{
#define AL(i, j) Al[(i) * ((i) + 1) / 2 + (j)]
#define LL(i, j) Lm[(i) * ((i) + 1) / 2 + (j)]
    // one reciprocal square root per column instead of a square root and up to six divisions: the solve is the
    // longest dependent chain of an LM iteration
    double Lm[21], inv[6], tr = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) tr += AL(i, i);
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double d = AL(j, j);
#pragma unroll
        for (int k = 0; k < j; k++) d -= LL(j, k) * LL(j, k);
        if (!(d > tr * 1e-13)) ok = false;
        const double rs = rsqrt(d);
        inv[j] = rs;
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            double s = AL(i, j);
#pragma unroll
            for (int k = 0; k < j; k++) s -= LL(i, k) * LL(j, k);
            LL(i, j) = s * rs;
        }
    }
    if (!ok) return false;
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double s = b[i];
#pragma unroll
        for (int k = 0; k < i; k++) s -= LL(i, k) * y[k];
        y[i] = s * inv[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        double s = y[i];
#pragma unroll
        for (int k = i + 1; k < 6; k++) s -= LL(k, i) * x[k];
        x[i] = s * inv[i];
    }
    return true;
#undef AL
#undef LL
}

// damping factors 10^k, k = -16 .. 16 (the dependency evaluates exp(k ln 10); the difference is a few ulp of lambda)
__constant__ double c_pow10[33] = {1e-16, 1e-15, 1e-14, 1e-13, 1e-12, 1e-11, 1e-10, 1e-9, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 1e0,
                                   1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16};

__device__ __noinline__ void sym_solve6_fallback(const double *Al, const double *b, double *x)
{
    double A[36];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j <= i; j++) A[i * 6 + j] = A[j * 6 + i] = Al[i * (i + 1) / 2 + j];
    sym_solve6(A, b, x);
}

// planar-homography initial pose from the 4 normalised points (all lanes of a group compute it redundantly)
__device__ __noinline__ void pnp_init_planar(const double *obj_xy /* 8 */, const double *mn /* 8 */, double *param)
{
    double A[64], b[8];
    for (int i = 0; i < 64; i++) A[i] = 0;
    for (int i = 0; i < 4; i++) {
        double X = obj_xy[2 * i], Y = obj_xy[2 * i + 1], x = mn[2 * i], y = mn[2 * i + 1];
        double *r0 = A + i * 8, *r1 = A + (i + 4) * 8;
        r0[0] = X; r0[1] = Y; r0[2] = 1; r0[6] = -x * X; r0[7] = -x * Y; b[i] = x;
        r1[3] = X; r1[4] = Y; r1[5] = 1; r1[6] = -y * X; r1[7] = -y * Y; b[i + 4] = y;
    }
    for (int a = 0; a < 6; a++) param[a] = 0;
    if (lu_solve8(A, b)) {
        const double H[9] = {b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7], 1.};
        double h1n = sqrt(H[0] * H[0] + H[3] * H[3] + H[6] * H[6]);
        double h2n = sqrt(H[1] * H[1] + H[4] * H[4] + H[7] * H[7]);
        double s1 = 1. / fmax(h1n, DBL_EPSILON), s2 = 1. / fmax(h2n, DBL_EPSILON), st = 2. / fmax(h1n + h2n, DBL_EPSILON);
        double h1[3] = {H[0] * s1, H[3] * s1, H[6] * s1}, h2[3] = {H[1] * s2, H[4] * s2, H[7] * s2};
        double h3[3] = {h1[1] * h2[2] - h1[2] * h2[1], h1[2] * h2[0] - h1[0] * h2[2], h1[0] * h2[1] - h1[1] * h2[0]};
        double R0[9] = {h1[0], h2[0], h3[0], h1[1], h2[1], h3[1], h1[2], h2[2], h3[2]};
        double r[3], R[9];
        rodrigues_mat2vec(R0, r);
        rodrigues_vec2mat(r, R, nullptr);
        rodrigues_mat2vec(R, r);
        param[0] = r[0]; param[1] = r[1]; param[2] = r[2];
        param[3] = H[2] * st; param[4] = H[5] * st; param[5] = H[8] * st;
    }
}

// all 4 lanes of the group call this together; j = lane's corner, (u, v) its image point, h = half marker length
__device__ void solve_pnp_planar_g4(unsigned gmask, int gbase, int j, double h, double u, double v, const CamModel &C,
                                    double *rvec_out, double *tvec_out)
{
    const double *k = C.k;
    const double X = (j == 0 || j == 3) ? -h : h, Y = (j < 2) ? h : -h;
    double mx, my;
    {   // undistortPoints, exactly 5 iterations, own corner
        const double ifx = 1. / C.fx, ify = 1. / C.fy;
        double x = (u - C.cx) * ifx, y = (v - C.cy) * ify, x0 = x, y0 = y;
        for (int it = 0; it < 5; it++) {
            double r2 = x * x + y * y;
            double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0) { x = (u - C.cx) * ifx; y = (v - C.cy) * ify; break; }
            double dX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
            double dY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
            x = (x0 - dX) * icdist;
            y = (y0 - dY) * icdist;
        }
        mx = x; my = y;
    }
    double param[6];
    {
        double oxy[8], mn[8];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            oxy[2 * q] = (q == 0 || q == 3) ? -h : h; oxy[2 * q + 1] = (q < 2) ? h : -h;
            mn[2 * q] = __shfl_sync(gmask, mx, gbase + q); mn[2 * q + 1] = __shfl_sync(gmask, my, gbase + q);
        }
        pnp_init_planar(oxy, mn, param);
    }
    double prev[6], JtJ[21], JtErr[6], R[9], dRdr[27];
    double prevErrNorm = DBL_MAX, errNorm = 0;
    int lambdaLg10 = -3, iters = 0;
    for (;;) {
        double pu, pv, dr[6], dt[6];
        rodrigues_reg<true>(param, R, dRdr);
        project_one<true>(X, Y, 0., R, dRdr, param + 3, C, pu, pv, dr, dt);
        const double e0 = pu - u, e1 = pv - v;
        double j0[6] = {dr[0], dr[1], dr[2], dt[0], dt[1], dt[2]}, j1[6] = {dr[3], dr[4], dr[5], dt[3], dt[4], dt[5]};
#pragma unroll
        for (int a = 0; a < 6; a++) {
#pragma unroll
            for (int b = 0; b <= a; b++) JtJ[a * (a + 1) / 2 + b] = gsum4(gmask, j0[a] * j0[b] + j1[a] * j1[b]);
            JtErr[a] = gsum4(gmask, j0[a] * e0 + j1[a] * e1);
        }
#pragma unroll
        for (int a = 0; a < 6; a++) prev[a] = param[a];
        if (iters == 0) prevErrNorm = sqrt(gsum4(gmask, e0 * e0 + e1 * e1));
        for (;;) {
            double A[21], d[6], lambda = c_pow10[min(max(lambdaLg10, -16), 16) + 16];
#pragma unroll
            for (int i = 0; i < 21; i++) A[i] = JtJ[i];
#pragma unroll
            for (int a = 0; a < 6; a++) A[a * (a + 1) / 2 + a] *= 1. + lambda;
            if (!chol_solve6_reg(A, JtErr, d)) sym_solve6_fallback(A, JtErr, d);
#pragma unroll
            for (int a = 0; a < 6; a++) param[a] = prev[a] - d[a];
            rodrigues_reg<false>(param, R, nullptr);
            project_one<false>(X, Y, 0., R, nullptr, param + 3, C, pu, pv, nullptr, nullptr);
            const double f0 = pu - u, f1 = pv - v;
            errNorm = sqrt(gsum4(gmask, f0 * f0 + f1 * f1));
            if (errNorm > prevErrNorm && ++lambdaLg10 <= 16) continue;
            break;
        }
        lambdaLg10 = max(lambdaLg10 - 1, -16);
        double nd = 0, np = 0;
#pragma unroll
        for (int a = 0; a < 6; a++) { double dd = param[a] - prev[a]; nd += dd * dd; np += prev[a] * prev[a]; }
        if (++iters >= 20 || sqrt(nd) / sqrt(np) < FLT_EPSILON) break;
        prevErrNorm = errNorm;
    }
    if (j == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++) { rvec_out[a] = param[a]; tvec_out[a] = param[3 + a]; }
    }
}

__global__ void __launch_bounds__(128) k_pose4(const float *__restrict__ corners, int n, const float *__restrict__ marker_len,
                                               float marker_len_all, CamModel C, double *__restrict__ rvec, double *__restrict__ tvec)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x, i = t >> 2, j = t & 3, lane = threadIdx.x & 31;
    if (i >= n) return;
    const float L = marker_len ? marker_len[i] : marker_len_all;
    const float hh = L / 2.f;  // legacy API: float marker length, float32 object points
    const unsigned gmask = 0xfu << (lane & ~3);
    solve_pnp_planar_g4(gmask, lane & ~3, j, (double)hh, (double)corners[8 * (size_t)i + 2 * j], (double)corners[8 * (size_t)i + 2 * j + 1], C,
                        rvec + 3 * (size_t)i, tvec + 3 * (size_t)i);
}

__global__ void __launch_bounds__(128) k_pose_frames4(const float *__restrict__ corners, const int32_t *__restrict__ n_markers, int batch,
                                                      int max_markers, const float *__restrict__ marker_len, float marker_len_all, CamModel C,
                                                      double *__restrict__ rvec, double *__restrict__ tvec)
{
    // work item = marker slot (f, m), compacted over the frames' marker counts would need a scan; slots beyond
    // n_markers[f] exit at once (4 lanes of a slot together)
    const int t = blockIdx.x * blockDim.x + threadIdx.x, i = t >> 2, j = t & 3, lane = threadIdx.x & 31;
    if (i >= batch * max_markers) return;
    const int f = i / max_markers, m = i - f * max_markers;
    if (m >= n_markers[f]) return;
    const float L = marker_len ? marker_len[f] : marker_len_all;
    const float hh = L / 2.f;
    const unsigned gmask = 0xfu << (lane & ~3);
    solve_pnp_planar_g4(gmask, lane & ~3, j, (double)hh, (double)corners[8 * (size_t)i + 2 * j], (double)corners[8 * (size_t)i + 2 * j + 1], C,
                        rvec + 3 * (size_t)i, tvec + 3 * (size_t)i);
}

__global__ void k_project_points(const double *__restrict__ obj, int n, const double *__restrict__ rvec,
                                 const double *__restrict__ tvec, CamModel C, double *__restrict__ img)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r[3] = {rvec[0], rvec[1], rvec[2]}, t[3] = {tvec[0], tvec[1], tvec[2]};
    project_points(obj + 3 * (size_t)i, 1, r, t, C, img + 2 * (size_t)i, nullptr, nullptr);
}

__global__ void k_project_points_multi(const double *__restrict__ obj, int n, const int32_t *__restrict__ pose_idx,
                                       const double *__restrict__ rvecs, const double *__restrict__ tvecs, CamModel C,
                                       double *__restrict__ img)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = pose_idx[i];
    double r[3] = {rvecs[3 * p], rvecs[3 * p + 1], rvecs[3 * p + 2]}, t[3] = {tvecs[3 * p], tvecs[3 * p + 1], tvecs[3 * p + 2]};
    project_points(obj + 3 * (size_t)i, 1, r, t, C, img + 2 * (size_t)i, nullptr, nullptr);
}

static int make_cam(apse_ctx *ctx, const double K[9], const double D[14], CamModel *C)
{
    if (!K || !D) CTX_FAIL(ctx, APSE_ERR_INVALID_ARG, "camera matrix / distortion missing");
    if (D[12] != 0 || D[13] != 0) CTX_FAIL(ctx, APSE_ERR_UNSUPPORTED, "tilted sensor model (tauX/tauY) is not supported");
    C->fx = K[0]; C->fy = K[4]; C->cx = K[2]; C->cy = K[5];
    for (int i = 0; i < 12; i++) C->k[i] = D[i];
    return APSE_OK;
}

int apse_pose(apse_ctx *ctx, const float *corners, int n, const float *marker_len, float marker_len_all, const double K[9],
              const double D[14], double *rvec, double *tvec, void *stream)
{
    if (!ctx || !corners || !rvec || !tvec || n < 0) CTX_FAIL(ctx, APSE_ERR_INVALID_ARG, "pose: bad argument");
    if (n == 0) return APSE_OK;
    CamModel C;
    int rc = make_cam(ctx, K, D, &C);
    if (rc) return rc;
    KLAUNCH(ctx, KID_POSE, (cudaStream_t)stream, k_pose4<<<div_up(4 * n, 128), 128, 0, (cudaStream_t)stream>>>(corners, n, marker_len, marker_len_all, C, rvec, tvec));
    return APSE_OK;
}

int apse_project_points(apse_ctx *ctx, const double *obj, int n, const double *rvec, const double *tvec, const double K[9],
                        const double D[14], double *img, void *stream)
{
    if (!ctx || !obj || !rvec || !tvec || !img || n <= 0) CTX_FAIL(ctx, APSE_ERR_INVALID_ARG, "project_points: bad argument");
    CamModel C;
    int rc = make_cam(ctx, K, D, &C);
    if (rc) return rc;
    KLAUNCH(ctx, KID_PROJECT, (cudaStream_t)stream, k_project_points<<<div_up(n, 64), 64, 0, (cudaStream_t)stream>>>(obj, n, rvec, tvec, C, img));
    return APSE_OK;
}

int apse_pose_frames(apse_ctx *ctx, const float *corners, const int32_t *n_markers, int batch, int max_markers,
                     const float *marker_len, float marker_len_all, const double K[9], const double D[14], double *rvec,
                     double *tvec, void *stream)
{
    if (!ctx || !corners || !n_markers || !rvec || !tvec || batch <= 0 || max_markers <= 0)
        CTX_FAIL(ctx, APSE_ERR_INVALID_ARG, "pose_frames: bad argument");
    CamModel C;
    int rc = make_cam(ctx, K, D, &C);
    if (rc) return rc;
    KLAUNCH(ctx, KID_POSE, (cudaStream_t)stream, k_pose_frames4<<<div_up(4 * batch * max_markers, 128), 128, 0, (cudaStream_t)stream>>>(corners, n_markers, batch, max_markers, marker_len,
                                                                               marker_len_all, C, rvec, tvec));
    return APSE_OK;
}

int apse_project_points_multi(apse_ctx *ctx, const double *obj, int n, const int32_t *pose_idx, const double *rvecs,
                              const double *tvecs, const double K[9], const double D[14], double *img, void *stream)
{
    if (!ctx || !obj || !pose_idx || !rvecs || !tvecs || !img || n <= 0) CTX_FAIL(ctx, APSE_ERR_INVALID_ARG, "project_points_multi: bad argument");
    CamModel C;
    int rc = make_cam(ctx, K, D, &C);
    if (rc) return rc;
    KLAUNCH(ctx, KID_PROJECT, (cudaStream_t)stream, k_project_points_multi<<<div_up(n, 64), 64, 0, (cudaStream_t)stream>>>(obj, n, pose_idx, rvecs, tvecs, C, img));
    return APSE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// aruco_detect.py:352-358  LED read-out: sum of the (2 half + 1)^2 gray neighbourhood `gray[y-half:y+half+1,
// x-half:x+half+1]` of projected points, with the slicing rules of the reference's numpy expression (a negative
// start wraps around, the stop is clamped: near the top / left edge the slice is empty and the sum is 0, near the
// bottom / right edge it is truncated).  One warp per point; pts = (frame, x, y) triples.
__device__ __forceinline__ void np_slice(int start, int stop, int len, int &a, int &b)
{
    if (start < 0) { start += len; if (start < 0) start = 0; }
    if (stop < 0) { stop += len; if (stop < 0) stop = 0; }
    a = min(start, len); b = min(stop, len);
}

__global__ void k_patch_sums(const uint8_t *__restrict__ gray, int w, int h, const int32_t *__restrict__ pts, int n, int half,
                             long long *__restrict__ sums)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n) return;
    const int f = pts[3 * i], x = pts[3 * i + 1], y = pts[3 * i + 2];
    int x0, x1, y0, y1;
    np_slice(x - half, x + half + 1, w, x0, x1);
    np_slice(y - half, y + half + 1, h, y0, y1);
    const int pw = max(x1 - x0, 0), ph = max(y1 - y0, 0);
    const uint8_t *g = gray + (size_t)f * w * h;
    long long s = 0;
    for (int k = lane; k < pw * ph; k += 32) s += g[(size_t)(y0 + k / pw) * w + x0 + k % pw];
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) sums[i] = s;
}

int apse_patch_sums(apse_ctx *ctx, const uint8_t *gray, int w, int h, const int32_t *pts, int n, int half, int64_t *sums, void *stream)
{
    if (!ctx || !gray || !pts || !sums || n <= 0 || half < 0 || w <= 0 || h <= 0) CTX_FAIL(ctx, APSE_ERR_INVALID_ARG, "patch_sums: bad argument");
    KLAUNCH(ctx, KID_PROJECT, (cudaStream_t)stream, k_patch_sums<<<div_up(n * 32, 128), 128, 0, (cudaStream_t)stream>>>(gray, w, h, pts, n, half, (long long *)sums));
    return APSE_OK;
}
