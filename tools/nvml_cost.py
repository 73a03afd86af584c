"""Development aid: how much does polling NVML perturb the pipeline? (bench.py samples clocks during the timed region)"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, pynvml
import apse_uav_b200 as A
from apse_uav_b200 import aruco
import __graft_entry__ as G
import bench
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
B = 60
K, D = bench.load_camera()
d = aruco.getPredefinedDictionary(aruco.DICT_4X4_50)
base = torch.from_numpy(bench.base_frames(6)).cuda()
seq = torch.stack([torch.roll(base[k % 6], shifts=(k % 7, k % 11), dims=(0, 1)) for k in range(2 * B)]).reshape(2, B, 2160, 3840, 3)
pipe = A.Pipeline(K, D, (3840, 2160), G.gamma_lut(), d, G.reference_parameters(aruco), max_batch=B, max_markers=64, streams=3)
def run(mode, period):
    stop = [False]; lat = []
    def poll():
        while not stop[0]:
            t0 = time.perf_counter()
            if mode in ("clock", "both"): pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            if mode in ("reasons", "both"): pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            if mode == "sleep": pass
            lat.append(time.perf_counter() - t0)
            time.sleep(period)
    th = threading.Thread(target=poll, daemon=True)
    for i in range(3): pipe.run(seq[i % 2])
    torch.cuda.synchronize()
    if mode != "none": th.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(30): pipe.run(seq[i % 2])
    b.record(); torch.cuda.synchronize(); stop[0] = True
    print(f"mode={mode} period={period}: {a.elapsed_time(b) / 30:.3f} ms/step; polls {len(lat)} mean latency {1e3 * np.mean(lat) if lat else 0:.2f} ms", flush=True)
for mode, period in (("none", 0), ("sleep", 0.05), ("clock", 0.05), ("reasons", 0.05), ("both", 0.05), ("both", 0.2), ("none", 0)):
    run(mode, period)
