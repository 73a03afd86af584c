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


def lab_tables():
    g = np.empty(256, np.uint16)
    c = np.empty(3072, np.uint16)
    ly = np.empty(256, np.uint16)
    lf = np.empty(256, np.uint16)
    ig = np.empty(4096, np.uint8)
    u16 = lambda a: _p(a, C.c_uint16)
    lib().orc_lab_tables(u16(g), u16(c), u16(ly), u16(lf), u8p(ig))
    return dict(gamma=g, cbrt=c, ly=ly, lf=lf, invgamma=ig)
