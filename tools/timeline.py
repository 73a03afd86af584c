"""Launch timeline of the overlapped (multi-stream) bench loop: every kernel launch of every context bracketed by CUDA
events, times relative to one process-wide origin (apse_timing_trace).  Development aid; the events perturb the
overlap slightly, so the figures explain the headline number, they are not a bench value.

    python tools/timeline.py [--batch 60] [--streams 3] [--steps 4] [--out gpurun_out/timeline.json]
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=60)
    ap.add_argument("--streams", type=int, default=3)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.json"))
    args = ap.parse_args()
    import torch
    import apse_uav_b200 as A
    from apse_uav_b200 import aruco
    import __graft_entry__ as G
    import bench

    K, D = bench.load_camera()
    d = aruco.getPredefinedDictionary(aruco.DICT_4X4_50)
    pipe = A.Pipeline(K, D, (bench.W, bench.H), G.gamma_lut(), d, G.reference_parameters(aruco), max_batch=args.batch,
                      max_markers=64, streams=args.streams, ring=args.steps + 4)
    base = torch.from_numpy(bench.base_frames(6)).cuda()
    seq = torch.stack([torch.roll(base[j % 6], shifts=(j % 7, j % 5), dims=(0, 1)) for j in range(args.batch)])
    seq2 = torch.roll(seq, shifts=(3, 9), dims=(1, 2))
    torch.cuda.synchronize()
    for i in range(3):
        A.Pipeline.wait(pipe.run_batch(seq if i % 2 else seq2, sync=False, input_ready=True))
    torch.cuda.synchronize()
    lib = pipe.engine.lib
    for e in pipe.engines:
        lib.apse_timing_trace(e.h, None, 4096)
        e.timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dets = [pipe.run_batch(seq if i % 2 else seq2, sync=False, input_ready=True) for i in range(args.steps)]
    for det in dets:
        A.Pipeline.wait(det)
    e1.record()
    torch.cuda.synchronize()
    rows = []
    for s, e in enumerate(pipe.engines):
        buf = (C.c_double * (3 * 4096))()
        n = lib.apse_timing_trace(e.h, buf, 4096)
        a = np.frombuffer(buf, dtype=np.float64)[:3 * n].reshape(n, 3)
        for kid, t0, t1 in a:
            rows.append((s, lib.apse_kernel_name(int(kid)).decode(), float(t0), float(t1)))
        e.timing(False)
    t_origin = min(r[2] for r in rows)
    rows = sorted((s, k, t0 - t_origin, t1 - t_origin) for s, k, t0, t1 in rows)
    rows.sort(key=lambda r: r[2])
    total = e0.elapsed_time(e1)
    print(f"{args.steps} steps x {args.batch} frames, {args.streams} streams: {total:.3f} ms ({total / args.steps:.3f} ms/step)")
    for s, k, t0, t1 in rows:
        print(f"ctx{s} {k:22s} {t0:9.3f} -> {t1:9.3f}  ({t1 - t0:7.3f} ms)")
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"total_ms": total, "steps": args.steps, "batch": args.batch, "streams": args.streams,
               "rows": rows}, open(args.out, "w"))


if __name__ == "__main__":
    main()
