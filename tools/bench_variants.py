"""Development aid: ms/step of Pipeline.run on HBM-resident frames for several stream counts, with and without the
per-kernel event timing hooks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import apse_uav_b200 as A
from apse_uav_b200 import aruco
import __graft_entry__ as G
import bench

B = 60
K, D = bench.load_camera()
d = aruco.getPredefinedDictionary(aruco.DICT_4X4_50)
base = torch.from_numpy(bench.base_frames(6)).cuda()
seq = torch.stack([torch.roll(base[k % 6], shifts=(k % 7, k % 11), dims=(0, 1)) for k in range(2 * B)]).reshape(2, B, 2160, 3840, 3)
for streams in (1, 2, 3, 4):
    pipe = A.Pipeline(K, D, (3840, 2160), G.gamma_lut(), d, G.reference_parameters(aruco), max_batch=B, max_markers=64, streams=streams)
    for timing in (False, True):
        for e in pipe.engines: e.timing(timing)
        for i in range(3): pipe.run(seq[i % 2])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record()
        for i in range(12): det = pipe.run(seq[i % 2])
        b.record(); cpu = time.perf_counter() - t0; torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 12
        print(f"streams={streams} timing={timing}: {ms:.3f} ms/step ({1e3 * B / ms:.0f} frames/s), cpu issue {1e3 * cpu / 12:.3f} ms/step", flush=True)
        for e in pipe.engines: e.timing_collect(reset=True); e.timing(False)
    pipe.close(); del pipe
    torch.cuda.empty_cache()
