"""First-contact GPU check (development aid): stage-by-stage diffs of the CUDA path against the oracle
(and cv2 when importable) plus rough stage timings.  Run on the GPU box: python tools/gpu_check.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import apse_uav_b200 as A
from apse_uav_b200 import aruco
from oracle import oracle as O
from tools import synth

cam = json.load(open("tests/golden/cam_params.json"))
K = np.array(cam["mtx"]); D = np.array(cam["dist"]).ravel()
W, H = 3840, 2160
lut = np.array([min(255, int(((i / 255.0) ** 2) * 255.0)) for i in range(256)], np.uint8)
d = aruco.getPredefinedDictionary(aruco.DICT_4X4_50)
p = aruco.DetectorParameters()
p.minMarkerPerimeterRate = 0.01; p.perspectiveRemovePixelPerCell = 8; p.perspectiveRemoveIgnoredMarginPerCell = 0.33
p.errorCorrectionRate = 2.0; p.aprilTagMinClusterPixels = 100; p.aprilTagMaxNmaxima = 5
p.aprilTagCriticalRad = 20 * np.pi / 180; p.aprilTagMaxLineFitMse = 1; p.aprilTagMinWhiteBlackDiff = 100
p.cornerRefinementMethod = aruco.CORNER_REFINE_APRILTAG

B = int(os.environ.get("B", "8"))
pipe = A.Pipeline(K, D, (W, H), lut, d, p, max_batch=B, max_markers=256)
e = pipe.engine
print("engine up; launches", e.launches)

# ---- maps
omx, omy = O.init_undistort_map(K, D, W, H)
mx, my = e.init_undistort_map(K, D, W, H)
print("map mismatches vs oracle:", int((mx.cpu().numpy() != omx).sum()), int((my.cpu().numpy() != omy).sum()))

frames = [synth.make_frame(d.bytesList, 3), synth.make_dense_frame(d.bytesList, 11)]
for fi, f in enumerate(frames):
    ref_bgr, ref_gray = O.preprocess(f, omx, omy, lut)
    t = torch.from_numpy(f).cuda()
    out, gray = e.preprocess(t, want_bgr=True)
    print(f"[frame {fi}] preprocess mismatches bgr {int((out.cpu().numpy() != ref_bgr).sum())} gray {int((gray.cpu().numpy() != ref_gray).sum())}")
    # stand-alone kernels
    r1 = e.remap(t, mx, my).cpu().numpy(); o1 = O.remap(f, omx, omy)
    l1 = e.cvt(torch.from_numpy(o1).cuda(), "rgb2lab").cpu().numpy(); ol = O.rgb2lab(o1)
    b1 = e.cvt(torch.from_numpy(ol).cuda(), "lab2rgb").cpu().numpy(); ob = O.lab2rgb(ol)
    print(f"   remap {int((r1 != o1).sum())} rgb2lab {int((l1 != ol).sum())} lab2rgb {int((b1 != ob).sum())}")
    # apriltag taps
    class P: pass
    op = O.AtParams(p.aprilTagMinClusterPixels, p.aprilTagMaxNmaxima, p.aprilTagCriticalRad, p.aprilTagMaxLineFitMse, p.aprilTagMinWhiteBlackDiff)
    oq, st = O.at_quads(ref_gray, op, dumps=True)
    e.set_params(p); e.set_dictionary(d.raw, d.markerSize, d.maxCorrectionBits)
    dbg = e.debug_apriltag(torch.from_numpy(ref_gray).cuda())
    th = dbg["thresh"].cpu().numpy(); lab = dbg["labels"].cpu().numpy().view(np.uint32)
    m = th != 127
    print(f"   thresh mismatches {int((th != st['thresh']).sum())}; label mismatches (non-127) {int((lab[m] != st['rep'][m]).sum())} of {int(m.sum())}")
    print(f"   points {dbg['points']} vs {st['points']}; clusters {dbg['clusters']} vs {st['clusters']}; fitted {dbg['fitted']} vs {st['fitted']}; quads {dbg['n_quads']} vs {len(oq)}")
    gq = dbg["quads"].cpu().numpy().reshape(-1, 8); oq8 = oq.reshape(-1, 8)
    if len(gq) and len(oq8):
        dist = np.abs(gq[:, None, :] - oq8[None, :, :]).max(-1)
        print(f"   quads: exact matches {int((dist.min(1) == 0).sum())}/{len(gq)}; max nearest diff {dist.min(1).max():.3g}; unmatched oracle quads {int((dist.min(0) > 1e-3).sum())}")
    # full detect
    oc, oi, orj = O.detect_markers_apriltag(ref_gray, d.raw, p)
    c, ids, rej = aruco.detectMarkers(ref_gray, d, parameters=p)
    ids_ = ids.ravel() if ids is not None else np.zeros(0, int)
    cc = np.array(c).reshape(-1, 4, 2)
    print(f"   detect: n {len(ids_)} vs {len(oi)}; ids equal {np.array_equal(ids_, oi)}; corner maxdiff {np.abs(cc - oc).max() if cc.shape == oc.shape and len(cc) else None}; rejected {len(rej)} vs {len(orj)}")
    # pose
    if len(ids_):
        rv, tv, _ = aruco.estimatePoseSingleMarkers(c, 0.55, K, D)
        orv, otv = O.estimate_pose_single_markers(cc, 0.55, K, D)
        rr = np.linalg.norm(rv - orv, axis=-1) / np.linalg.norm(orv, axis=-1); rt = np.linalg.norm(tv - otv, axis=-1) / np.linalg.norm(otv, axis=-1)
        print(f"   pose rel diff: rvec max {rr.max():.3g} tvec max {rt.max():.3g}")
    try:
        import cv2
        pc = cv2.aruco.DetectorParameters()
        for k in vars(p):
            setattr(pc, k, getattr(p, k))
        rc, ri, rrj = cv2.aruco.ArucoDetector(cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_4X4_50), pc).detectMarkers(ref_gray)
        ri_ = ri.ravel() if ri is not None else np.zeros(0, int)
        print(f"   vs cv2: ids equal {np.array_equal(ids_, ri_)}; corner maxdiff {np.abs(cc - np.array(rc).reshape(-1,4,2)).max() if len(ri_) == len(ids_) and len(ids_) else None}")
    except ImportError:
        print("   cv2 not importable here")

# ---- timing (device-resident batch)
batch = torch.from_numpy(np.stack([frames[i % 2] for i in range(B)])).cuda()
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n): fn()
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / n
gray = e.preprocess(batch)[1]
print(f"timing B={B}: preprocess {timed(lambda: e.preprocess(batch)):.3f} ms; preprocess+bgr {timed(lambda: e.preprocess(batch, want_bgr=True)):.3f} ms; "
      f"detect {timed(lambda: e.detect(gray, max_markers=256, want_rejected=False)):.3f} ms; full {timed(lambda: pipe.run_batch(batch)):.3f} ms")
sparse = torch.from_numpy(np.stack([frames[0]] * B)).cuda()
gs = e.preprocess(sparse)[1]
print(f"timing sparse B={B}: detect {timed(lambda: e.detect(gs, max_markers=256, want_rejected=False)):.3f} ms; full {timed(lambda: pipe.run_batch(sparse)):.3f} ms")
res = pipe.to_host(pipe.run_batch(batch))
print("n per frame", res["n"].tolist(), "launches", e.launches)
