"""Development aid: where the 3 ms of the native sequence post-pass go (1800 frames, one GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from apse_uav_b200 import sequence, shard
import apse_uav_b200 as A


class Args: batch = 60; max_markers = 64; streams = 3
pipe = bench.make_pipeline(Args, 0)
base = torch.from_numpy(bench.base_sequence(12)).cuda()
plan = bench.sequence_plan(1800, 12)
frames = torch.stack([torch.roll(base[p], shifts=(dy, dx), dims=(0, 1)) for p, dy, dx in plan[:600]])
det = pipe.run_sequence(frames)
det = {k: torch.cat([v, v, v], 0) for k, v in det.items()}   # 1800 frames of results
torch.cuda.synchronize()
e = pipe.engine
orig = {n: getattr(sequence, n) for n in ("_host_copy", "scan", "run_jobs", "finish")}
acc = {}


def timed(name):
    f = orig[name]

    def g(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize(); acc[name] = acc.get(name, 0) + time.perf_counter() - t0
        return r
    return g


for n in orig:
    setattr(sequence, n, timed(n))
pf = e.pose_frames


def pose_t(*a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = pf(*a, **k); torch.cuda.synchronize(); acc["pose_frames"] = acc.get("pose_frames", 0) + time.perf_counter() - t0; return r


e.pose_frames = pose_t
for rep in range(3):
    acc.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rows = sequence.postpass_device(e, det)
    t1 = time.perf_counter()
    csv = sequence.rows_to_csv(rows)
    t2 = time.perf_counter()
    print("postpass %.2f ms, csv %.2f ms | " % (1e3 * (t1 - t0), 1e3 * (t2 - t1)) + ", ".join("%s %.2f" % (k, 1e3 * v) for k, v in acc.items()))
