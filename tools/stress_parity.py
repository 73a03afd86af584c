"""Parity stress: N seeded synthetic 4K frames (sparse + every 6th dense) through the 3-stream overlapped Pipeline, each
frame compared with the CPU oracle chain (ids, order, corners <= 1e-3 px, pose <= 1e-4 rel).  Repeats the GPU run to check
run-to-run determinism."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import apse_uav_b200 as A
from apse_uav_b200 import aruco
import __graft_entry__ as G
from oracle import oracle as O
from tools import synth

N = int(os.environ.get("N", "24"))
K, D = bench.load_camera()
d = aruco.getPredefinedDictionary(aruco.DICT_4X4_50)
p = G.reference_parameters(aruco)
lut = G.gamma_lut()
frames = []
for i in range(N):
    if i % 6 == 5:
        frames.append(synth.make_dense_frame(d.bytesList, 500 + i))
    else:
        frames.append(synth.make_frame(d.bytesList, 7000 + 13 * i, bench.W, bench.H, ids=(1, 2, 3, 4, 11, 23)[: 4 + i % 3], side_range=(40, 110)))
frames = np.stack(frames)
pipe = A.Pipeline(K, D, (bench.W, bench.H), lut, d, p, max_batch=N, max_markers=256, streams=3, ring=4)
dev = torch.from_numpy(frames).cuda()
runs = []
for rep in range(3):
    det = pipe.run_batch(dev, want_rejected=True, sync=False, input_ready=True)
    runs.append(A.Pipeline.to_host(det))
for k in ("n", "ids", "corners", "n_rejected", "rejected", "rvec", "tvec"):
    for r in runs[1:]:
        assert np.array_equal(r[k], runs[0][k]), ("non-deterministic", k)
res = runs[0]
mx, my = O.init_undistort_map(K, D, bench.W, bench.H)
bad = 0
t0 = time.time()
for i in range(N):
    _, gray = O.preprocess(frames[i], mx, my, lut)
    oc, oi, orj = O.detect_markers_apriltag(gray, d.raw, p)
    n = int(res["n"][i])
    ok = n == len(oi) and np.array_equal(res["ids"][i, :n], oi) and (n == 0 or np.abs(res["corners"][i, :n] - oc).max() <= 1e-3)
    ok = ok and int(res["n_rejected"][i]) == len(orj)
    if ok and n:
        orv, otv = O.estimate_pose_single_markers(oc, 0.55, K, D)
        rel = np.linalg.norm(res["tvec"][i, :n] - otv[:, 0], axis=-1) / np.linalg.norm(otv[:, 0], axis=-1)
        ok = rel.max() < 1e-4
    print(f"frame {i}: markers {n} (oracle {len(oi)}), rejected {int(res['n_rejected'][i])} (oracle {len(orj)}) -> {'ok' if ok else 'MISMATCH'}", flush=True)
    bad += not ok
print(f"{N} frames, {bad} mismatches, deterministic over 3 runs, oracle time {time.time() - t0:.1f} s")
sys.exit(1 if bad else 0)
