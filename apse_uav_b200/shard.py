"""Frame-wise sharding of a sequence across GPUs (one process + one Pipeline per GPU, no collective on the hot
path) and the two-pass handling of the only cross-frame coupling, the marker-length
recurrence of aruco_detect.py:306-308,601,623,641.

Pass 1 (parallel, per rank): preprocess + detect + pose with the nominal marker length on the rank's block of frames.
Gather  (off the hot path): the per-frame results (a few KB per frame) go to rank 0 (gather_object).
Scan    (rank 0, sequential): the post-pass over frames in order with tvec scaled by L_k / L_nominal (tvec is
        linear in the marker length), which yields the marker length L_k each frame's pose must use.
Pass 2 (rank 0, one batched launch): exact pose of every frame with its L_k, then the final sequential scan.
"""
from __future__ import annotations

import numpy as np

from .postpass import SequencePostPass, MARKER_LENGTH_ORG, csv_line


def shard_bounds(n_frames: int, world: int):
    """Contiguous blocks: rank r gets frames [lo, hi)."""
    return [(n_frames * r // world, n_frames * (r + 1) // world) for r in range(world)]


def pack_results(host, lo):
    """Per-frame records from Pipeline.to_host output."""
    recs = []
    for i in range(len(host["n"])):
        n = int(host["n"][i])
        recs.append(dict(frame=lo + i, ids=host["ids"][i, :n].copy(), corners=host["corners"][i, :n].copy(),
                         rvec=host["rvec"][i, :n].copy(), tvec=host["tvec"][i, :n].copy()))
    return recs


def gather_records(local, rank=0, world=1, group=None):
    """All ranks' records on rank 0, ordered by frame (small python objects; works on gloo and nccl groups)."""
    if world == 1:
        return sorted(local, key=lambda r: r["frame"])
    import torch.distributed as dist
    out = [None] * world if rank == 0 else None
    dist.gather_object(local, out, dst=0, group=group)
    if rank != 0:
        return None
    return sorted([r for part in out for r in part], key=lambda r: r["frame"])


def scan_marker_lengths(records, project, start_frame=1):
    """First sequential scan: poses were computed with MARKER_LENGTH_ORG; rescale tvec by L_k / L_nominal."""
    pp = SequencePostPass(project, start_frame=start_frame)
    lengths = []
    for r in records:
        L = pp.marker_length
        lengths.append(L)
        pp.step(start_frame + r["frame"], r["ids"] if len(r["ids"]) else None, r["corners"], r["rvec"],
                r["tvec"] * (L / MARKER_LENGTH_ORG))
    return lengths


def exact_pose(records, lengths, pose_fn):
    """Second pass: pose_fn(corners (n,4,2) f32, marker_len (n,) f32) -> (rvec (n,3), tvec (n,3)) in ONE batched call."""
    counts = [len(r["ids"]) for r in records]
    if sum(counts) == 0:
        return records
    corners = np.concatenate([r["corners"].reshape(-1, 4, 2) for r in records if len(r["ids"])]).astype(np.float32)
    ml = np.concatenate([np.full(c, L, np.float32) for c, L in zip(counts, lengths) if c])
    rv, tv = pose_fn(corners, ml)
    out, o = [], 0
    for r, c in zip(records, counts):
        out.append(dict(r, rvec=rv[o:o + c], tvec=tv[o:o + c]))
        o += c
    return out


def final_scan(records, project, start_frame=1, led_mean_for_frame=None, led_sums_for_frame=None):
    pp = SequencePostPass(project, start_frame=start_frame)
    rows = []
    for r in records:
        if led_mean_for_frame is not None:
            pp.led_mean = led_mean_for_frame(r["frame"])
        if led_sums_for_frame is not None:
            pp.led_sums = led_sums_for_frame(r["frame"])
        rows.append(pp.step(start_frame + r["frame"], r["ids"] if len(r["ids"]) else None, r["corners"], r["rvec"], r["tvec"]))
    return rows


def run_sequence(pipe, frames, rank=0, world=1, group=None, start_frame=1, leds=False):
    """frames: this rank's block of the sequence ([n_local,H,W,3] uint8 CUDA tensor); lo = first global index of the
    block is derived from shard_bounds over the total length exchanged below.  Returns CSV rows on rank 0.
    leds=True (single process): the corrected gray frames stay on the device and the LED strip of the host vehicle
    (aruco_detect.py:338-373) is read back by the GPU patch-sum kernel in the final scan."""
    if leds and world > 1:
        raise ValueError("LED read-out needs the gray frames of every rank on rank 0: run it per rank (world = 1)")
    n_local = int(frames.shape[0])
    if world > 1:
        import torch.distributed as dist
        sizes = [None] * world
        dist.all_gather_object(sizes, n_local, group=group)
        lo = sum(sizes[:rank])
    else:
        lo = 0
    det = pipe.run(frames, want_gray=leds) if n_local else None
    gray = det.pop("gray") if (leds and det is not None) else None
    host = pipe.to_host(det) if n_local else dict(n=np.zeros(0, np.int32))
    records = gather_records(pack_results(host, lo) if n_local else [], rank, world, group)
    if rank != 0:
        return None
    e = pipe.engine
    project = lambda obj, rvec, tvec: e.project_points(obj, rvec, tvec).cpu().numpy()

    def pose_fn(corners, ml):
        rv, tv = e.pose(corners, ml)
        return rv.cpu().numpy(), tv.cpu().numpy()

    lengths = scan_marker_lengths(records, project, start_frame)
    records = exact_pose(records, lengths, pose_fn)
    led_fn = None
    if gray is not None:
        led_fn = lambda frame: (lambda xy: e.patch_sums(gray, [(frame, int(x), int(y)) for x, y in xy], half=2))
    return final_scan(records, project, start_frame, led_sums_for_frame=led_fn)


def rows_to_csv(rows):
    from .postpass import CSV_HEADER
    return "\n".join([CSV_HEADER] + [csv_line(r) for r in rows]) + "\n"
