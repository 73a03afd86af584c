"""Development aid: TMA-staged fused preprocess vs the oracle on 4K frames (+ timing of the kernel alone)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from apse_uav_b200.engine import Engine
from apse_uav_b200 import aruco
from oracle import oracle as O
from tools import synth
import __graft_entry__ as G

cam = json.load(open("tests/golden/cam_params.json"))
K = np.array(cam["mtx"]); D = np.array(cam["dist"]).ravel()
W, H = 3840, 2160
B = int(os.environ.get("B", "8"))
lut = G.gamma_lut()
d = aruco.getPredefinedDictionary(aruco.DICT_4X4_50)
e = Engine(0, W, H, B)
e.set_camera(K, D, W, H); e.set_lut(lut)
frames = np.stack([synth.make_frame(d.bytesList, 3 + i) if i % 2 == 0 else synth.make_dense_frame(d.bytesList, 11 + i) for i in range(min(B, 3))])
t = torch.from_numpy(np.stack([frames[i % len(frames)] for i in range(B)])).cuda()
out, gray = e.preprocess(t, want_bgr=True)
torch.cuda.synchronize()
_, gray2 = e.preprocess(t, want_bgr=False)
torch.cuda.synchronize()
print("launched ok")
if os.environ.get("CHECK", "1") == "1":
    omx, omy = O.init_undistort_map(K, D, W, H)
    for i in range(len(frames)):
        rb, rg = O.preprocess(frames[i], omx, omy, lut)
        print(f"frame {i}: bgr mismatches {int((out[i].cpu().numpy() != rb).sum())} gray {int((gray[i].cpu().numpy() != rg).sum())} gray(no bgr) {int((gray2[i].cpu().numpy() != rg).sum())}")
    print("all frames equal to their base:", all(torch.equal(gray[i], gray[i % len(frames)]) for i in range(B)))
for want in (False, True):
    for _ in range(3): e.preprocess(t, want_bgr=want)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): e.preprocess(t, want_bgr=want)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"B={B} want_bgr={want}: {ms:.3f} ms/launch = {ms / B * 1e3:.1f} us/frame -> {33177600 * B / ms / 1e6:.0f} GB/s algorithmic")
