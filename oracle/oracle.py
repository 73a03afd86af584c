"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the C oracle (oracle/*.c -> oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import this module.
The product package never does (tests/test_boundary.py greps for it).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


u8p = lambda a: _p(a, C.c_uint8)
f32p = lambda a: _p(a, C.c_float)
f64p = lambda a: _p(a, C.c_double)
i32p = lambda a: _p(a, C.c_int32)
u32p = lambda a: _p(a, C.c_uint32)


# ------------------------------------------------------------------------------------------------- stage 1
def init_undistort_map(K, D, w, h):
    K = np.ascontiguousarray(K, np.float64).ravel()
    Df = np.zeros(14, np.float64)
    Df[:np.size(D)] = np.asarray(D, np.float64).ravel()
    mx = np.empty((h, w), np.float32)
    my = np.empty((h, w), np.float32)
    lib().orc_init_undistort_map(f64p(K), f64p(Df), w, h, f32p(mx), f32p(my))
    return mx, my


def remap(src, mapx, mapy):
    src = np.ascontiguousarray(src)
    cn = 1 if src.ndim == 2 else src.shape[2]
    h, w = mapx.shape
    dst = np.empty((h, w) + (() if src.ndim == 2 else (cn,)), np.uint8)
    lib().orc_remap_bilinear(u8p(src), src.shape[1], src.shape[0], cn, f32p(np.ascontiguousarray(mapx)),
                             f32p(np.ascontiguousarray(mapy)), w, h, u8p(dst))
    return dst


def undistort(src, K, D, newK=None):
    """cv2.undistort(src, K, D, None, newK) for 8-bit 1- or 3-channel images (FP64 coordinate rounded to Q5 directly)"""
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape[:2]
    cn = 1 if src.ndim == 2 else src.shape[2]
    dst = np.empty_like(src)
    Kc = np.ascontiguousarray(K, np.float64).ravel()
    Dc = _k14(D)
    Nc = np.ascontiguousarray(newK, np.float64).ravel() if newK is not None else None
    dp = lambda a: _p(a, C.c_double)
    lib().orc_undistort(u8p(src), w, h, cn, dp(Kc), dp(Dc), dp(Nc) if Nc is not None else None, u8p(dst))
    return dst


def rgb2lab(src):
    src = np.ascontiguousarray(src)
    dst = np.empty_like(src)
    lib().orc_rgb2lab_u8(u8p(src), C.c_size_t(src.size // 3), u8p(dst))
    return dst


def lab2rgb(src):
    src = np.ascontiguousarray(src)
    dst = np.empty_like(src)
    lib().orc_lab2rgb_u8(u8p(src), C.c_size_t(src.size // 3), u8p(dst))
    return dst


def bgr2gray(src):
    src = np.ascontiguousarray(src)
    dst = np.empty(src.shape[:-1], np.uint8)
    lib().orc_bgr2gray(u8p(src), C.c_size_t(src.size // 3), u8p(dst))
    return dst


def preprocess(bgr, mapx, mapy, lut):
    bgr = np.ascontiguousarray(bgr)
    h, w = bgr.shape[:2]
    out = np.empty_like(bgr)
    gray = np.empty((h, w), np.uint8)
    lut = np.ascontiguousarray(lut, np.uint8).ravel()
    lib().orc_preprocess(u8p(bgr), w, h, f32p(np.ascontiguousarray(mapx)), f32p(np.ascontiguousarray(mapy)),
                         u8p(lut), u8p(out), u8p(gray))
    return out, gray


def quad_image(gray, decimate=0.0, sigma=0.0):
    """Image of the APRILTAG quad detector for aprilTagQuadDecimate / aprilTagQuadSigma (aruco_detect.py:203,231-233):
    (image, corner scale)."""
    q = np.ascontiguousarray(gray)
    scale = np.float32(1.0)
    if decimate > 1:
        h, w = q.shape
        dw, dh = C.c_int(0), C.c_int(0)
        lib().orc_resize_area_dsize(w, h, C.c_float(decimate), C.byref(dw), C.byref(dh))
        dw, dh = dw.value, dh.value
        out = np.empty((dh, dw), np.uint8)
        scale_f = 1.0 / float(np.float32(1.0) / np.float32(decimate))   # 1 / fx, fx = 1.f / decimate in float32
        if abs(scale_f - round(scale_f)) < 2.220446049250313e-16:    # integer factor: the dependency's block-average fast path
            lib().orc_resize_area_int(u8p(q), w, h, int(round(scale_f)), dw, dh, u8p(out))
        else:
            lib().orc_resize_area(u8p(q), w, h, dw, dh, C.c_double(scale_f), u8p(out))
        q, scale = out, np.float32(decimate)
    if sigma != 0:
        out = np.empty_like(q)
        if lib().orc_quad_sigma(u8p(q), q.shape[1], q.shape[0], C.c_float(sigma), u8p(out)) != 0:
            raise ValueError("aprilTagQuadSigma too large")
        q = out
    return q, scale


def lab_tables():
    g = np.empty(256, np.uint16)
    c = np.empty(3072, np.uint16)
    ly = np.empty(256, np.uint16)
    lf = np.empty(256, np.uint16)
    ig = np.empty(4096, np.uint8)
    u16 = lambda a: _p(a, C.c_uint16)
    lib().orc_lab_tables(u16(g), u16(c), u16(ly), u16(lf), u8p(ig))
    return dict(gamma=g, cbrt=c, ly=ly, lf=lf, invgamma=ig)


# ------------------------------------------------------------------------------------------- APRILTAG path
class AtParams(C.Structure):
    _fields_ = [("min_cluster_pixels", C.c_int), ("max_nmaxima", C.c_int), ("critical_rad", C.c_float),
                ("max_line_fit_mse", C.c_float), ("min_white_black_diff", C.c_int)]

    @classmethod
    def from_cv(cls, p):
        return cls(int(p.aprilTagMinClusterPixels), int(p.aprilTagMaxNmaxima), float(p.aprilTagCriticalRad),
                   float(p.aprilTagMaxLineFitMse), int(p.aprilTagMinWhiteBlackDiff))


def fast_atan2(y, x):
    f = lib().orc_fast_atan2
    f.restype = C.c_float
    f.argtypes = [C.c_float, C.c_float]
    return f(y, x)


def at_threshold(gray, min_wb_diff):
    gray = np.ascontiguousarray(gray)
    out = np.empty_like(gray)
    lib().orc_at_threshold(u8p(gray), gray.shape[1], gray.shape[0], int(min_wb_diff), u8p(out))
    return out


def at_unionfind(thresh):
    thresh = np.ascontiguousarray(thresh)
    rep = np.empty(thresh.shape, np.uint32)
    lib().orc_at_unionfind(u8p(thresh), thresh.shape[1], thresh.shape[0], u32p(rep))
    return rep


def at_quads(gray, params, max_quads=4096, dumps=False):
    """Raw quads of the AprilTag-style detector, (n,4,2) float32, dependency candidate order."""
    gray = np.ascontiguousarray(gray)
    h, w = gray.shape
    P = params if isinstance(params, AtParams) else AtParams.from_cv(params)
    q = np.zeros((max_quads, 8), np.float32)
    stats = np.zeros(3, np.int64)
    t = np.empty((h, w), np.uint8) if dumps else None
    r = np.empty((h, w), np.uint32) if dumps else None
    n = lib().orc_at_quads(u8p(gray), w, h, C.byref(P), f32p(q), max_quads, _p(stats, C.c_int64),
                           u8p(t) if dumps else None, u32p(r) if dumps else None)
    if n > max_quads:
        raise RuntimeError("oracle quad capacity exceeded")
    res = q[:n].reshape(n, 4, 2).copy()
    if dumps:
        return res, dict(points=int(stats[0]), clusters=int(stats[1]), fitted=int(stats[2]), thresh=t, rep=r)
    return res


# --------------------------------------------------------------------------- candidate filter + decoding
class DecParams(C.Structure):
    _fields_ = [("marker_size", C.c_int), ("border_bits", C.c_int), ("cell_size", C.c_int),
                ("cell_margin_rate", C.c_double), ("min_otsu_stddev", C.c_double),
                ("max_border_err_rate", C.c_double), ("error_correction_rate", C.c_double),
                ("max_correction_bits", C.c_int), ("min_distance_to_border", C.c_int),
                ("min_marker_distance_rate", C.c_double), ("min_group_distance", C.c_float),
                ("detect_inverted", C.c_int), ("skip_decoded_parents", C.c_int)]

    @classmethod
    def from_cv(cls, p, marker_size=4, max_correction_bits=1, skip_decoded_parents=0):
        return cls(marker_size, int(p.markerBorderBits), int(p.perspectiveRemovePixelPerCell),
                   float(p.perspectiveRemoveIgnoredMarginPerCell), float(p.minOtsuStdDev),
                   float(p.maxErroneousBitsInBorderRate), float(p.errorCorrectionRate), max_correction_bits,
                   int(p.minDistanceToBorder), float(p.minMarkerDistanceRate), float(p.minGroupDistance),
                   int(bool(p.detectInvertedMarker)), skip_decoded_parents)


def perspective_transform(src, dst):
    M = np.empty(9, np.float64)
    lib().orc_perspective_transform(f32p(np.ascontiguousarray(src, np.float32)),
                                    f32p(np.ascontiguousarray(dst, np.float32)), f64p(M))
    return M.reshape(3, 3)


def warp_nearest(gray, corners, S):
    gray = np.ascontiguousarray(gray)
    out = np.empty((S, S), np.uint8)
    lib().orc_warp_nearest(u8p(gray), gray.shape[1], gray.shape[0],
                           f32p(np.ascontiguousarray(corners, np.float32)), S, u8p(out))
    return out


def otsu(px):
    px = np.ascontiguousarray(px, np.uint8)
    return lib().orc_otsu(u8p(px), px.size)


def extract_bits(gray, corners, dp):
    gray = np.ascontiguousarray(gray)
    n = dp.marker_size + 2 * dp.border_bits
    bits = np.empty((n, n), np.uint8)
    thr = C.c_int(0)
    lib().orc_extract_bits(u8p(gray), gray.shape[1], gray.shape[0], f32p(np.ascontiguousarray(corners, np.float32)),
                           C.byref(dp), u8p(bits), None, C.byref(thr))
    return bits, thr.value


def identify(inner_bits, bytes_list, max_corr_bits, rate):
    inner_bits = np.ascontiguousarray(inner_bits, np.uint8)
    bl = np.ascontiguousarray(bytes_list, np.uint8)
    idx, rot = C.c_int(-1), C.c_int(0)
    ok = lib().orc_identify(u8p(inner_bits), inner_bits.shape[0], u8p(bl), bl.shape[0], int(max_corr_bits),
                            C.c_double(rate), C.byref(idx), C.byref(rot))
    return bool(ok), idx.value, rot.value


def identify_candidates(gray, quads, dp, bytes_list, max_out=4096):
    gray = np.ascontiguousarray(gray)
    quads = np.ascontiguousarray(quads, np.float32).reshape(-1, 8)
    bl = np.ascontiguousarray(bytes_list, np.uint8)
    corners = np.zeros((max_out, 8), np.float32)
    rejected = np.zeros((max_out, 8), np.float32)
    ids = np.zeros(max_out, np.int32)
    nrej = C.c_int(0)
    na = lib().orc_identify_candidates(u8p(gray), gray.shape[1], gray.shape[0], f32p(quads), len(quads),
                                       C.byref(dp), u8p(bl), bl.shape[0], f32p(corners), i32p(ids),
                                       f32p(rejected), max_out, C.byref(nrej))
    return corners[:na].reshape(na, 4, 2).copy(), ids[:na].copy(), rejected[:nrej.value].reshape(-1, 4, 2).copy()


def detect_markers_apriltag(gray, bytes_list, cvparams, marker_size=4, max_correction_bits=1):
    """aruco_detect.py:267 (APRILTAG mode) -> (corners (n,4,2) f32, ids (n,) i32, rejected (m,4,2) f32)."""
    dec, sigma = float(getattr(cvparams, "aprilTagQuadDecimate", 0.0)), float(getattr(cvparams, "aprilTagQuadSigma", 0.0))
    if dec > 1 or sigma != 0:   # quads on the shrunk / blurred image, scaled back in float32; identification on the original
        qim, scale = quad_image(gray, dec, sigma)
        quads = (at_quads(qim, cvparams) * scale).astype(np.float32)
    else:
        quads = at_quads(gray, cvparams)
    dp = DecParams.from_cv(cvparams, marker_size, max_correction_bits)
    return identify_candidates(gray, quads, dp, bytes_list)


# ------------------------------------------------------------------------------------------------- pose
def _k14(D):
    k = np.zeros(14, np.float64)
    k[:np.size(D)] = np.asarray(D, np.float64).ravel()
    return k


def project_points(obj, rvec, tvec, K, D, jacobian=False):
    obj = np.ascontiguousarray(np.asarray(obj, np.float64).reshape(-1, 3))
    n = len(obj)
    img = np.empty((n, 2), np.float64)
    K = np.ascontiguousarray(K, np.float64).ravel()
    k = _k14(D)
    r = np.ascontiguousarray(np.asarray(rvec, np.float64).ravel())
    t = np.ascontiguousarray(np.asarray(tvec, np.float64).ravel())
    if jacobian:
        dr = np.empty((2 * n, 3), np.float64)
        dt = np.empty((2 * n, 3), np.float64)
        lib().orc_project_points(f64p(obj), n, f64p(r), f64p(t), f64p(K), f64p(k), f64p(img), f64p(dr), f64p(dt))
        return img, dr, dt
    lib().orc_project_points(f64p(obj), n, f64p(r), f64p(t), f64p(K), f64p(k), f64p(img), None, None)
    return img


def undistort_points(pts, K, D):
    pts = np.ascontiguousarray(np.asarray(pts, np.float64).reshape(-1, 2))
    out = np.empty_like(pts)
    lib().orc_undistort_points(f64p(pts), len(pts), f64p(np.ascontiguousarray(K, np.float64).ravel()),
                               f64p(_k14(D)), f64p(out))
    return out


def rodrigues(r):
    R = np.empty(9, np.float64)
    J = np.empty(27, np.float64)
    lib().orc_rodrigues(f64p(np.ascontiguousarray(r, np.float64).ravel()), f64p(R), f64p(J))
    return R.reshape(3, 3), J.reshape(3, 9)


def rodrigues_inv(R):
    r = np.empty(3, np.float64)
    lib().orc_rodrigues_inv(f64p(np.ascontiguousarray(R, np.float64).ravel()), f64p(r))
    return r


def estimate_pose_single_markers(corners, marker_length, K, D):
    """aruco_detect.py:601 -> (rvecs (n,1,3) f64, tvecs (n,1,3) f64)."""
    c = np.ascontiguousarray(np.asarray(corners, np.float32).reshape(-1, 8))
    n = len(c)
    rv = np.zeros((n, 1, 3), np.float64)
    tv = np.zeros((n, 1, 3), np.float64)
    lib().orc_estimate_pose_single_markers(f32p(c), n, C.c_float(marker_length),
                                           f64p(np.ascontiguousarray(K, np.float64).ravel()), f64p(_k14(D)),
                                           f64p(rv), f64p(tv))
    return rv, tv


# -------------------------------------------------------------------------------------------- classic path
class ClassicParams(C.Structure):
    _fields_ = [("win_min", C.c_int), ("win_max", C.c_int), ("win_step", C.c_int), ("constant", C.c_double),
                ("min_perimeter_rate", C.c_double), ("max_perimeter_rate", C.c_double),
                ("approx_accuracy_rate", C.c_double), ("min_corner_distance_rate", C.c_double),
                ("min_distance_to_border", C.c_int)]

    @classmethod
    def from_cv(cls, p):
        return cls(int(p.adaptiveThreshWinSizeMin), int(p.adaptiveThreshWinSizeMax), int(p.adaptiveThreshWinSizeStep),
                   float(p.adaptiveThreshConstant), float(p.minMarkerPerimeterRate), float(p.maxMarkerPerimeterRate),
                   float(p.polygonalApproxAccuracyRate), float(p.minCornerDistanceRate), int(p.minDistanceToBorder))


def adaptive_threshold(gray, win, c):
    """cv2.adaptiveThreshold(gray, 255, ADAPTIVE_THRESH_MEAN_C, THRESH_BINARY_INV, win, c) (even win -> win + 1)."""
    gray = np.ascontiguousarray(gray)
    out = np.empty_like(gray)
    lib().orc_adaptive_threshold(u8p(gray), gray.shape[1], gray.shape[0], int(win), C.c_double(c), u8p(out))
    return out


def find_contours(binary):
    """cv2.findContours(binary, RETR_LIST, CHAIN_APPROX_NONE) -> list of (n,2) int32 arrays in cv2's order."""
    b = np.ascontiguousarray(binary, np.uint8)
    h, w = b.shape
    max_pts = 2 * w * h + 16
    max_c = w * h // 2 + 16
    pts = np.empty((max_pts, 2), np.int32)
    off = np.empty(max_c + 1, np.int64)
    n = lib().orc_find_contours(u8p(b), w, h, i32p(pts), C.c_longlong(max_pts), _p(off, C.c_int64), max_c)
    if n < 0:
        raise RuntimeError("oracle contour capacity exceeded")
    return [pts[off[i]:off[i + 1]].copy() for i in range(n)]


def approx_poly_dp(contour, eps):
    c = np.ascontiguousarray(np.asarray(contour, np.int32).reshape(-1, 2))
    out = np.empty((len(c) + 4, 2), np.int32)
    n = lib().orc_approx_poly_dp(i32p(c), len(c), C.c_double(eps), i32p(out), len(out))
    return out[:n].copy()


def is_contour_convex(poly):
    p = np.ascontiguousarray(np.asarray(poly, np.int32).reshape(-1, 2))
    return bool(lib().orc_is_contour_convex(i32p(p), len(p)))


def classic_quads(gray, cvparams, max_quads=1 << 16, refined=False):
    """Candidate quads of the classic path, (n,4,2) float32, dependency order (windows ascending, contour order).
    refined=True: also the CORNER_REFINE_CONTOUR corners of every candidate (same corner order)."""
    gray = np.ascontiguousarray(gray)
    P = cvparams if isinstance(cvparams, ClassicParams) else ClassicParams.from_cv(cvparams)
    q = np.zeros((max_quads, 8), np.float32)
    r = np.zeros((max_quads, 8), np.float32) if refined else None
    n = lib().orc_classic_quads_refined(u8p(gray), gray.shape[1], gray.shape[0], C.byref(P), f32p(q), f32p(r) if refined else None, max_quads)
    if n > max_quads:
        raise RuntimeError("oracle quad capacity exceeded")
    if refined:
        return q[:n].reshape(n, 4, 2).copy(), r[:n].reshape(n, 4, 2).copy()
    return q[:n].reshape(n, 4, 2).copy()


def refine_candidate_lines(contour, corners):
    c = np.ascontiguousarray(np.asarray(contour, np.int32).reshape(-1, 2))
    out = np.zeros(8, np.float32)
    rc = lib().orc_refine_candidate_lines(i32p(c), len(c), f32p(np.ascontiguousarray(corners, np.float32).ravel()), f32p(out))
    return out.reshape(4, 2) if rc == 0 else None


def corner_subpix(gray, corners, win, max_iter, eps):
    gray = np.ascontiguousarray(gray)
    c = np.ascontiguousarray(np.asarray(corners, np.float32).reshape(-1, 2)).copy()
    lib().orc_corner_subpix(u8p(gray), gray.shape[1], gray.shape[0], f32p(c), len(c), int(win), int(max_iter),
                            C.c_double(eps))
    return c


def detect_markers_classic(gray, bytes_list, cvparams, marker_size=4, max_correction_bits=1):
    """aruco_detect.py:267 with cornerRefinementMethod NONE (0), SUBPIX (1) or CONTOUR (2) -> (corners, ids, rejected)."""
    method = int(cvparams.cornerRefinementMethod)
    if method == 2:
        quads, refined = classic_quads(gray, cvparams, refined=True)
    else:
        quads = classic_quads(gray, cvparams)
    dp = DecParams.from_cv(cvparams, marker_size, max_correction_bits)
    corners, ids, rejected = identify_candidates(gray, quads, dp, bytes_list)
    if method == 2:
        # every accepted marker takes the refined corners of its candidate: the first candidate (dependency order) with the
        # same corners up to the rotation the identification applied
        for i in range(len(corners)):
            done = False
            for qi in range(len(quads)):
                for sh in range(4):
                    if np.array_equal(np.roll(quads[qi], -sh, axis=0), corners[i]):
                        corners[i] = np.roll(refined[qi], -sh, axis=0)
                        done = True
                        break
                if done:
                    break
    if int(cvparams.cornerRefinementMethod) == 1 and len(corners):
        nb = marker_size + 2 * int(cvparams.markerBorderBits)
        for i in range(len(corners)):
            c = corners[i]
            side = np.float32(0)
            for a in range(4):
                b = (a + 1) % 4
                dx, dy = np.float32(c[a, 0] - c[b, 0]), np.float32(c[a, 1] - c[b, 1])
                side = np.float32(side + np.sqrt(np.float32(dx * dx + dy * dy)))
            module = np.float32(side / np.float32(4.0 * nb))
            win = max(1, int(np.rint(np.float32(cvparams.relativeCornerRefinmentWinSize) * module)))
            win = min(win, int(cvparams.cornerRefinementWinSize))
            corners[i] = corner_subpix(gray, c, win, int(cvparams.cornerRefinementMaxIterations),
                                       float(cvparams.cornerRefinementMinAccuracy))
    return corners, ids, rejected
