"""Development aid: where does a pipeline step spend its time?  Per-phase CUDA-event times around the three library
calls of Pipeline.run_batch vs the per-kernel times recorded inside the library."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import apse_uav_b200 as A
from apse_uav_b200 import aruco
import __graft_entry__ as G
import bench

B = int(os.environ.get("B", "60")); dense = os.environ.get("DENSE", "0") == "1"
K, D = bench.load_camera()
d = aruco.getPredefinedDictionary(aruco.DICT_4X4_50)
pipe = A.Pipeline(K, D, (3840, 2160), G.gamma_lut(), d, G.reference_parameters(aruco), max_batch=B, max_markers=256 if dense else 64)
e = pipe.engine
S = int(os.environ.get("STREAMS", "1"))
if S > 1:
    pipe2 = A.Pipeline(K, D, (3840, 2160), G.gamma_lut(), d, G.reference_parameters(aruco), max_batch=B, max_markers=256 if dense else 64, streams=S)
from tools import synth
base = [synth.make_dense_frame(d.bytesList, 11 + i) if dense else synth.make_frame(d.bytesList, 1000 + i) for i in range(3)]
frames = torch.from_numpy(np.stack([base[i % 3] for i in range(B)])).cuda()
frames2 = torch.roll(frames, shifts=(3, 5), dims=(1, 2))
def ev(): return torch.cuda.Event(enable_timing=True)
for _ in range(3): pipe.run_batch(frames)
torch.cuda.synchronize()
for timing in (False, True):
    e.timing(timing); e.timing_collect(reset=True)
    N = 6
    t = [ev() for _ in range(4 * N + 1)]
    torch.cuda.synchronize(); w0 = time.perf_counter()
    t[0].record()
    for i in range(N):
        fr = frames if i % 2 == 0 else frames2
        _, gray = e.preprocess(fr); t[4 * i + 1].record()
        det = e.detect(gray, max_markers=pipe.max_markers, want_rejected=False); t[4 * i + 2].record()
        rv, tv = e.pose_frames(det["corners"], det["n"], 0.55); t[4 * i + 3].record()
        t[4 * i + 4].record()
    cpu_ms = 1e3 * (time.perf_counter() - w0)
    torch.cuda.synchronize()
    tot = t[0].elapsed_time(t[-1]) / N
    ph = [np.mean([t[4 * i + j].elapsed_time(t[4 * i + j + 1]) for i in range(N)]) for j in range(3)]
    kt = e.timing_collect(reset=True) if timing else {}
    print(f"timing={timing} B={B} dense={dense}: step {tot:.3f} ms (cpu issue {cpu_ms / N:.3f} ms/step)  preprocess {ph[0]:.3f}  detect {ph[1]:.3f}  pose {ph[2]:.3f}")
    if kt:
        print("   kernels (ms/step):", {k: round(v[0] / N, 3) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0])}, "sum", round(sum(v[0] for v in kt.values()) / N, 3))
e.timing(False)
if S > 1:
    for _ in range(3): pipe2.run_batch(frames)
    torch.cuda.synchronize()
    a, b = ev(), ev(); a.record()
    for i in range(6): out2 = pipe2.run_batch(frames if i % 2 == 0 else frames2)
    b.record(); torch.cuda.synchronize()
    print(f"streams={S}: step {a.elapsed_time(b) / 6:.3f} ms -> {1e3 * B * 6 / a.elapsed_time(b):.0f} frames/s; equal to 1-stream result:",
          torch.equal(out2["ids"], pipe.run_batch(frames2)["ids"]))
print("markers per frame:", det["n"][:6].tolist(), " -> frames/s", 1e3 * B / tot)
