"""Batched, device-resident marker pipeline: the per-frame work of aruco_detect.py:589-601
(preprocessFrame -> BGR2GRAY -> detectMarkers -> estimatePoseSingleMarkers) for many frames at once.

Frames are independent up to the pose scale, so a sequence is processed in batches on
one GPU and sharded frame-wise across GPUs (shard.py); the sequential marker-length / gating / distance logic of
aruco_detect.py:598-782 runs afterwards on the host over the small per-frame results (postpass.py).
"""
from __future__ import annotations

import numpy as np

from .engine import Engine
from ._lib import ApseError


class Pipeline:
    """`streams` > 1 splits every batch over that many CUDA streams, each with its own context (scratch), so that the
    latency-bound tail of one sub-batch (cluster scan, quad fit, decode, pose) overlaps the bandwidth/issue-bound
    preprocess of the next one."""

    def __init__(self, camera_matrix, dist_coeffs, size, lut, dictionary, params, max_batch=16, device=0,
                 max_markers=64, marker_length=0.55, streams=1):
        w, h = int(size[0]), int(size[1])
        streams = max(1, min(int(streams), max_batch))
        self.sub_batch = (max_batch + streams - 1) // streams
        bl = np.ascontiguousarray(dictionary.bytesList, np.uint8)
        self.engines = []
        for _ in range(streams):
            e = Engine(device, w, h, self.sub_batch)
            e.set_camera(camera_matrix, dist_coeffs, w, h)
            e.set_lut(lut)
            e.set_dictionary(bl.reshape(bl.shape[0], -1), dictionary.markerSize, dictionary.maxCorrectionBits)
            e.set_params(params)
            self.engines.append(e)
        self.engine = self.engines[0]
        torch = self.engine.torch
        self.streams = [torch.cuda.Stream(device=self.engine.tdev) for _ in range(streams)] if streams > 1 else []
        self.max_batch, self.max_markers, self.marker_length = max_batch, max_markers, float(marker_length)
        self.size = (w, h)

    @property
    def launches(self):
        return sum(e.launches for e in self.engines)

    def close(self):
        for e in self.engines:
            e.close()

    def _run_on(self, e, frames, out, want_gray, marker_length):
        gray = None
        if want_gray:
            gray = e.torch.empty(frames.shape[:3], dtype=e.torch.uint8, device=e.tdev)
        e.process_frames(frames, out, marker_length, gray=gray)   # one library call: K1t -> candidates -> decode -> pose
        return gray

    def run_batch(self, frames, want_gray=False, want_rejected=False, marker_length=None):
        """frames: [B,H,W,3] uint8 CUDA tensor, B <= max_batch.  Returns dict of device tensors."""
        e = self.engine
        torch = e.torch
        B = frames.shape[0]
        if B > self.max_batch:
            raise ApseError(-1, f"batch {B} exceeds max_batch {self.max_batch}")
        ml = self.marker_length if marker_length is None else marker_length
        det = e.alloc_detections(B, self.max_markers, want_rejected, pose=True)
        if not self.streams or B <= 1:
            gray = self._run_on(e, frames, det, want_gray, ml)
        else:
            cur = torch.cuda.current_stream(e.tdev)
            ready = cur.record_event()
            grays = []
            for s, (eng, st) in enumerate(zip(self.engines, self.streams)):
                lo, hi = s * self.sub_batch, min(B, (s + 1) * self.sub_batch)
                if lo >= hi:
                    break
                st.wait_event(ready)
                with torch.cuda.stream(st):
                    sl = {k: v[lo:hi] for k, v in det.items()}
                    mls = ml[lo:hi] if isinstance(ml, torch.Tensor) else ml
                    grays.append(self._run_on(eng, frames[lo:hi], sl, want_gray, mls))
                cur.wait_stream(st)
            gray = torch.cat(grays, 0) if want_gray else None
        if want_gray:
            det["gray"] = gray
        return det

    def run(self, frames, **kw):
        """Any number of frames, processed in batches of max_batch; results concatenated on the device."""
        torch = self.engine.torch
        outs = [self.run_batch(frames[i:i + self.max_batch], **kw) for i in range(0, frames.shape[0], self.max_batch)]
        if len(outs) == 1:
            return outs[0]
        return {k: torch.cat([o[k] for o in outs], 0) for k in outs[0]}

    @staticmethod
    def to_host(det):
        """Gather the small per-frame results to the host as numpy arrays (the only D2H traffic of the pipeline)."""
        out = {k: v.cpu().numpy() for k, v in det.items() if k != "gray"}
        bad = np.nonzero(out["status"])[0]
        if len(bad):
            raise ApseError(int(out["status"][bad[0]]), f"work-buffer capacity exceeded in frames {bad.tolist()[:8]}")
        return out
