"""Development aid: host-side cost of the native sequence post-pass (scan / finish / CSV text) on fabricated results of 1800 frames."""
import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from apse_uav_b200 import sequence, _lib
F, M = 1800, 4
rng = np.random.default_rng(0)
n = np.full(F, 4, np.int32)
ids = np.tile(np.array([4, 1, 2, 3], np.int32), (F, 1))
base = np.array([[100, 100], [200, 100], [200, 200], [100, 200]], np.float32)
corners = np.stack([np.stack([base + 300 * i + 0.01 * f for i in range(4)]) for f in range(F)]).astype(np.float32)
rvec = rng.normal(0, 0.1, (F, M, 3)); rvec[..., 0] += 3.0
tvec = rng.normal(0, 1, (F, M, 3)); tvec[..., 2] = 30 + rng.normal(0, 0.1, (F, M))
cfg = sequence.seq_config(1, 1, False, None, 3840, 2160)
for rep in range(3):
    t0 = time.perf_counter(); l, _, _ = sequence.scan(cfg, n, ids, corners, rvec, tvec, True, False)
    t1 = time.perf_counter(); l2, rows, jobs = sequence.scan(cfg, n, ids, corners, rvec, tvec, False, True)
    t2 = time.perf_counter()
    res = np.zeros(len(jobs), _lib.SEQ_RESULT_DTYPE)
    rows = sequence.finish(rows, res)
    t3 = time.perf_counter(); s = sequence.rows_to_csv(rows); t4 = time.perf_counter()
    print("scan1 %.3f scan2 %.3f finish %.3f csv %.3f ms; jobs %d, job bytes %d, csv %d B" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t4-t3), len(jobs), jobs.itemsize, len(s)))
import time
for name, fn in (("scan2", lambda: sequence.scan(cfg, n, ids, corners, rvec, tvec, False, True)), ("scan1", lambda: sequence.scan(cfg, n, ids, corners, rvec, tvec, True, False))):
    ts = []
    for _ in range(30):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    print(name, "median %.3f ms min %.3f" % (1e3 * sorted(ts)[15], 1e3 * min(ts)))
lib = _lib.load()
import ctypes as C
rows = np.zeros(F, sequence.SEQ_ROW_DTYPE); jobs = np.zeros(8000, _lib.SEQ_JOB_DTYPE); nj = C.c_int(0); lengths = np.empty(F)
ts = []
for _ in range(30):
    t0 = time.perf_counter()
    lib.apse_sequence_scan(C.byref(cfg), F, M, sequence._ptr(n), sequence._ptr(ids), sequence._ptr(corners), sequence._ptr(rvec), sequence._ptr(tvec), 0, sequence._ptr(lengths), sequence._ptr(rows), sequence._ptr(jobs), 8000, C.byref(nj))
    ts.append(time.perf_counter() - t0)
print("raw C scan2 median %.3f ms min %.3f" % (1e3 * sorted(ts)[15], 1e3 * min(ts)))
