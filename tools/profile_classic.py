"""Development aid: per-kernel times of the classic candidate path (BASELINE.json config 5) on dense 4K frames."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import apse_uav_b200 as A
from apse_uav_b200 import aruco
import __graft_entry__ as G
import bench
from tools import synth

B = int(os.environ.get("B", "8"))
K, D = bench.load_camera()
d = aruco.getPredefinedDictionary(aruco.DICT_4X4_50)
base = [synth.make_dense_frame(d.bytesList, 11 + i) for i in range(2)]
frames = torch.from_numpy(np.stack([base[i % 2] for i in range(B)])).cuda()
for wins in ((3, 23, 10), (13, 13, 1), (3, 53, 10)):
    p = G.reference_parameters(aruco); p.cornerRefinementMethod = 1
    p.adaptiveThreshWinSizeMin, p.adaptiveThreshWinSizeMax, p.adaptiveThreshWinSizeStep = wins
    pipe = A.Pipeline(K, D, (3840, 2160), G.gamma_lut(), d, p, max_batch=B, max_markers=512)
    e = pipe.engine
    for _ in range(2): det = pipe.run_batch(frames)
    torch.cuda.synchronize()
    e.timing(True); e.timing_collect(reset=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    N = 3
    a.record()
    for _ in range(N): det = pipe.run_batch(frames)
    b.record(); torch.cuda.synchronize()
    kt = e.timing_collect(reset=True); e.timing(False)
    ms = a.elapsed_time(b) / N
    print(f"classic wins={wins} B={B}: {ms:.2f} ms/batch = {ms / B:.2f} ms/frame -> {1e3 * B / ms:.0f} frames/s; markers {det['n'][:2].tolist()} status {det['status'][:2].tolist()}")
    print("   kernels (ms/batch):", {k: round(v[0] / N, 3) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0])})
    pipe.close()
