#!/usr/bin/env python
"""bench.py -- 4K frames/s of the marker pipeline (undistort + gamma -> detectMarkers(APRILTAG) -> pose).

    python bench.py --gpus N --steps K --warmup W          # our arm (one rank per GPU; torchrun for N > 1)
    python bench.py --impl reference --gpus N ...          # reference arm: cv2 on the host cores, rank 0 only

A step = one pass of the hot path over one batch of `--batch` synthetic 3840x2160 frames.  The default K x batch
= 30 x 60 = 1800 frames = BASELINE.json configs[2] (the 1800-frame sequence the metric is quoted on).
Timed region: barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks.
`value`   : frames already resident in HBM (each step reads a different 1.5 GB slice, far larger than L2).
`e2e`     : same metric through the public API with HOST (pinned) frames: H2D of the step's frames and D2H of
            its detections inside the timed region.
`roofline`: dominant kernel by device time (per-kernel CUDA events recorded inside libapse_b200), algorithmic
            bytes per launch from SURVEY.md 8(d) / DESIGN.md, peak from MEASURED_PEAKS.json.
`cpu_baseline`: the reference's own cv2 call chain (aruco_detect.py:250-269,592-601) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

W, H = 3840, 2160
FRAME_BYTES = W * H * 3
# algorithmic bytes per 4K frame (restated in DESIGN.md)
ALG_BYTES = {
    "k_preprocess_fused": 33_177_600,   # read BGR 24 883 200 + write gray 8 294 400
    "k_tile_minmax": 8_294_400,          # read gray (tile arrays negligible)
    "k_threshold": 16_588_800,           # read gray + write ternary
    "k_ccl_local": 41_472_000,           # read ternary + write labels (u32)
    "k_ccl_merge": 0, "k_ccl_flatten": 41_472_000, "k_emit_points": 41_472_000,
    "pipeline": 132_710_400,
}


def load_camera():
    cam = json.load(open(os.path.join(ROOT, "tests", "golden", "cam_params.json")))
    return np.array(cam["mtx"]), np.array(cam["dist"]).ravel()


def base_frames(n, seed=1000):
    """n distinct sparse frames (ids 1,2,3 vehicles + 4 host), seeded; host numpy uint8 [n,H,W,3]."""
    from tools import synth
    from apse_uav_b200 import aruco
    d = aruco.getPredefinedDictionary(aruco.DICT_4X4_50)
    return np.stack([synth.make_frame(d.bytesList, seed + i, W, H, ids=(1, 2, 3, 4), side_range=(50, 90)) for i in range(n)])


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons DURING the timed region, read in-process through NVML (nvidia-ml-py); spawning
    nvidia-smi inside the timed region stalls CUDA launches, so it is only the fallback for a single sample."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.nv, self.handle = index, [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _reasons(self):
        nv = self.nv
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        out = []
        for name, attr in (("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                           ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                           ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                           ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap")):
            bit = getattr(nv, attr, None)
            if bit is not None and mask & bit:
                out.append(name)
        return out

    def run(self):
        if self.nv is None:
            return
        while not self.stop_flag:
            try:
                self.samples.append((float(self.nv.nvmlDeviceGetClockInfo(self.handle, self.nv.NVML_CLOCK_SM)), self._reasons()))
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvml unavailable"}
        sm = [s[0] for s in self.samples]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": self.max_sm, "reasons": sorted({r for s in self.samples for r in s[1]}),
                "samples": len(sm), "source": "nvml, 50 ms period, during the timed region"}


# ---------------------------------------------------------------------------------------------------------------
class CpuReference:
    """The reference's cv2 call chain (aruco_detect.py:250-269,592-601 through the 4.13 compat shim)."""

    def __init__(self, threads, n_frames=4):
        import cv2
        from oracle import cv2_compat as C
        self.cv2, self.C, self.threads = cv2, C, threads
        self.K, self.D = load_camera()
        cv2.setNumThreads(threads)
        self.mapx, self.mapy = cv2.initUndistortRectifyMap(self.K, self.D, None, self.K, (W, H), 5)
        self.lut, self.params = C.gamma_lut(), C.reference_parameters()
        self.frames = base_frames(n_frames)
        self.markers = 0
        self.run(1)  # warm-up

    def run(self, n):
        for i in range(n):
            r = self.C.reference_chain(self.frames[i % len(self.frames)], self.mapx, self.mapy, self.lut, self.params,
                                       self.K, self.D)
            self.markers += 0 if r["ids"] is None else len(r["ids"])

    def info(self, n):
        return {"kind": "reference", "cores": self.threads,
                "sample": f"{n} synthetic 4K frames through cv2 {self.cv2.__version__} (remap+RGB2LAB+LUT+LAB2RGB+BGR2GRAY+"
                          f"ArucoDetector(APRILTAG)+solvePnP per marker), cv2.setNumThreads({self.threads})"}


def cpu_reference_fps(n_frames, threads):
    """Returns (frames/s, info dict) of the reference chain on the host cores; falls back to the C oracle port."""
    try:
        ref = CpuReference(threads, min(n_frames, 4))
    except ImportError:
        return port_fps(n_frames)
    t0 = time.perf_counter()
    ref.run(n_frames)
    dt = time.perf_counter() - t0
    return n_frames / dt, ref.info(n_frames)


def port_fps(n_frames, frames=None):
    import __graft_entry__  # noqa: F401  (sys.path)
    from oracle import oracle as O
    from apse_uav_b200 import aruco
    import __graft_entry__ as G
    K, D = load_camera()
    d = aruco.getPredefinedDictionary(aruco.DICT_4X4_50)
    p = G.reference_parameters(aruco)
    lut = G.gamma_lut()
    mx, my = O.init_undistort_map(K, D, W, H)
    if frames is None:
        frames = base_frames(min(n_frames, 2))
    t0 = time.perf_counter()
    for i in range(n_frames):
        _, gray = O.preprocess(frames[i % len(frames)], mx, my, lut)
        c, ids, _ = O.detect_markers_apriltag(gray, d.raw, p)
        O.estimate_pose_single_markers(c, 0.55, K, D)
    dt = time.perf_counter() - t0
    return n_frames / dt, {"kind": "port", "cores": 1, "sample": f"{n_frames} synthetic 4K frames through the C oracle (1 thread)"}


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = max(1, args.ref_frames_per_step)
    try:
        ref = CpuReference(threads, min(4, per_step))
        runner, info = ref.run, ref.info(args.steps * per_step)
    except ImportError:
        runner, info = (lambda n: port_fps(n)), {"kind": "port", "cores": 1, "sample": "C oracle, 1 thread"}
    for _ in range(args.warmup):
        runner(per_step)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        runner(per_step)
    dt = time.perf_counter() - t0
    fps = args.steps * per_step / dt
    line = {"impl": "reference", "metric": "4K frames/s (undistort+ArUco detect+pose)", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
            "config": workload_config(args, per_step),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": info["cores"], "kind": info["kind"],
                             "sample": f"{args.steps} steps x {per_step} frames; " + info["sample"]},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, frames_per_step):
    return {"workload": "configs[2]: full pipeline undistort+gamma -> detectMarkers(CORNER_REFINE_APRILTAG, DICT_4X4_50, "
                        "parameters of aruco_detect.py:190-203) -> estimatePoseSingleMarkers on a synthetic 3840x2160 "
                        "sequence (ids 1-4, sigma-3 noise)",
            "frames_per_step": frames_per_step, "frame": [W, H, 3],
            "sequence_frames": args.steps * frames_per_step,
            "l2": "every step reads a different slice of the HBM-resident sequence (>= 0.37 GB per step, L2 is 126 MB)",
            "streams_per_gpu": args.streams,
            "parallelism": f"frame-sharded x{args.gpus}, no collective on the hot path"}


# ---------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import apse_uav_b200 as A
    from apse_uav_b200 import aruco
    import __graft_entry__ as G

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, D = load_camera()
    d = aruco.getPredefinedDictionary(aruco.DICT_4X4_50)
    params = G.reference_parameters(aruco)
    Bt = args.batch
    chunk = min(Bt, 60)                 # frames per library call (context scratch is sized for this)
    pipe = A.Pipeline(K, D, (W, H), G.gamma_lut(), d, params, max_batch=chunk, device=local_rank, max_markers=args.max_markers,
                      streams=args.streams, ring=args.steps + args.warmup + 1)

    # ---- synthetic sequence resident in HBM: n_base seeded frames, each step is a distinct cyclic translation
    base = torch.from_numpy(base_frames(args.base_frames, seed=1000 + 97 * rank)).to(dev)
    n_slices = max(2, min(args.steps + args.warmup, int(args.hbm_gb * 1e9 // (Bt * FRAME_BYTES))))
    seq = torch.empty((n_slices, Bt, H, W, 3), dtype=torch.uint8, device=dev)
    g = torch.Generator().manual_seed(7 + rank)
    for s in range(n_slices):
        for j in range(Bt):
            k = s * Bt + j
            dx, dy = (int(v) for v in torch.randint(-40, 41, (2,), generator=g))
            seq[s, j] = torch.roll(base[k % len(base)], shifts=(dy, dx), dims=(0, 1))
    torch.cuda.synchronize()

    def step(i, serial=False, overlap=False):
        # overlap: the frames are resident and complete, and the result is only waited for at the end of the timed region,
        # so the chain of one batch runs under the preprocess of the next one (Pipeline docstring)
        return pipe.run_batch(seq[i % n_slices], serial=serial, sync=not overlap, input_ready=overlap)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        A.Pipeline.wait(step(i, overlap=True))
    # ---- timed region (device-resident frames): exactly K steps, CUDA events on the launching stream
    sampler = ClockSampler(local_rank)
    if not os.environ.get("APSE_BENCH_NOSAMPLER"):
        sampler.start()
    l0 = pipe.launches
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dets = []
    for i in range(args.steps):
        dets.append(step(args.warmup + i, overlap=True))   # results stay on the device until the region ends
    for d_ in dets:
        A.Pipeline.wait(d_)                                 # the current stream waits for every batch's chain ...
    e1.record()                                             # ... so this event closes the whole K-step region
    barrier()
    ms = e0.elapsed_time(e1)
    launches = pipe.launches - l0
    sampler.stop_flag = True
    markers = int(torch.stack([d_["n"] for d_ in dets]).sum().item())
    del dets

    # ---- per-kernel pass: the same K steps again with every launch bracketed by CUDA events inside the library
    # (apse_timing_*), sub-batches issued one after the other on ONE stream so that a kernel's event pair measures that
    # kernel alone (in the pass above kernels of different streams share the SMs and their durations overlap)
    for e in pipe.engines:
        e.timing(True)
    step(0, serial=True)
    for e in pipe.engines:
        e.timing_collect(reset=True)
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for i in range(args.steps):
        step(args.warmup + i, serial=True)
    k1.record()
    barrier()
    ms_instrumented = k0.elapsed_time(k1)
    ktimes = {}
    for e in pipe.engines:
        for k, (kms_, kc_) in e.timing_collect(reset=True).items():
            a_, b_ = ktimes.get(k, (0.0, 0))
            ktimes[k] = (a_ + kms_, b_ + kc_)
        e.timing(False)

    # ---- e2e: pinned host frames -> H2D -> pipeline -> D2H detections, per step inside the timed region
    n_host = min(args.steps, 3)
    host = [torch.empty((Bt, H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(n_host)]
    for j in range(n_host):
        host[j].copy_(seq[j % n_slices])
    d2h_bytes = 0

    def e2e_run(n):
        """n steps through the public host-frame API: H2D of every step's frames and D2H of its detections inside"""
        nonlocal d2h_bytes
        for out in pipe.run_host_stream(host[i % n_host] for i in range(n)):
            d2h_bytes = sum(v.nbytes for v in out.values())

    e2e_run(min(args.warmup, 2))
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0)

    # ---- reduce over ranks: max time, summed work
    stats = torch.tensor([ms, e2e_ms, float(launches), float(markers)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_ms, launches, markers = float(mx[0]), float(mx[1]), int(sm[2]), int(sm[3])
    if rank == 0:
        frames_total = args.steps * Bt * world
        value = frames_total / (ms / 1e3)
        e2e_value = e2e_steps * Bt * world / (e2e_ms / 1e3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        # dominant kernel by accumulated device time on rank 0
        dom = max(ktimes.items(), key=lambda kv: kv[1][0]) if ktimes else ("none", (0.0, 0))
        name, (kms, kcount) = dom
        frames_per_launch = args.steps * Bt / max(kcount, 1)
        alg = ALG_BYTES.get(name, ALG_BYTES["pipeline"]) * frames_per_launch
        achieved = alg / (kms / max(kcount, 1) / 1e3) / 1e9 if kms > 0 else 0.0
        ksum = sum(v[0] for v in ktimes.values())
        traffic, traffic_src = None, None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of the latest `ncu --set full` capture of this kernel, scaled per launch
            import glob
            tfile = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_k1_traffic.json")))[-1]
            tj = json.load(open(tfile))
            if name == "k_preprocess_fused":
                traffic = tj["dram_bytes_per_frame"] * frames_per_launch
                traffic_src = f"profiles/{os.path.basename(tfile)} (ncu --set full capture, bytes per frame x frames per launch)"
        except Exception:
            pass
        line = {
            "metric": "4K frames/s (undistort+ArUco detect+pose)", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32 pixels, f64 line fits and pose", "data": "synthetic",
            "config": workload_config(args, Bt),
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": Bt * FRAME_BYTES,
                    "d2h_bytes_per_step": int(d2h_bytes), "steps": e2e_steps},
            "gpu_launches": int(launches),
            "markers_found": markers,
            "pipeline_roofline": {"algorithmic_bytes_per_frame": ALG_BYTES["pipeline"],
                                  "achieved_gbs_per_gpu": ALG_BYTES["pipeline"] * value / world / 1e9,
                                  "frac_of_hbm_peak": ALG_BYTES["pipeline"] * value / world / 1e9 / peak},
            "roofline": {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                         "traffic_source": traffic_src,
                         "bound_note": "HBM is the roofline SURVEY.md 8(d) prescribes; the kernel itself is bound by SM issue (86 %), "
                                       "the shared-memory wavefront pipe (82 %) and the ALU pipe (76 %) -- "
                                       "profiles/r01e_k_preprocess_tma_ncu_full.csv -- not by DRAM (13 %)",
                         "share_of_kernel_time": kms / ksum if ksum else None,
                         "algorithmic_bytes_per_launch": alg, "launches": kcount, "avg_ms": kms / max(kcount, 1)},
            "kernel_ms_per_step": {k: v[0] / args.steps for k, v in sorted(ktimes.items(), key=lambda kv: -kv[1][0])},
            "kernel_timing_pass": {"ms_per_step": ms_instrumented / args.steps,
                                   "note": "same K steps repeated on one stream (sub-batches serialised) with a CUDA event pair "
                                           "around every launch; the headline pass overlaps sub-batches on several streams"},
            "clocks": sampler.summary(),
        }
        cpu_fps, info = cpu_reference_fps(args.cpu_frames, os.cpu_count() or 1) if world == 1 and args.cpu_frames > 0 else (None, None)
        if info:
            line["cpu_baseline"] = {"value": cpu_fps, "unit": "frames/s", "cores": info["cores"], "kind": info["kind"],
                                    "sample": info["sample"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=60, help="frames per step")
    ap.add_argument("--base-frames", type=int, default=6, help="distinct seeded frames rendered on the host")
    ap.add_argument("--hbm-gb", type=float, default=48.0, help="HBM budget of the resident synthetic sequence")
    ap.add_argument("--max-markers", type=int, default=64)
    ap.add_argument("--streams", type=int, default=3, help="CUDA streams (sub-batches in flight) per GPU")
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--cpu-frames", type=int, default=48, help="frames of the bounded CPU baseline sample (0 = skip)")
    ap.add_argument("--ref-frames-per-step", type=int, default=2)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # plain `python bench.py --gpus N`: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
