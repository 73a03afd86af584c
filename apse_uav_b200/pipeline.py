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
    def __init__(self, camera_matrix, dist_coeffs, size, lut, dictionary, params, max_batch=16, device=0,
                 max_markers=64, marker_length=0.55):
        w, h = int(size[0]), int(size[1])
        self.engine = Engine(device, w, h, max_batch)
        self.engine.set_camera(camera_matrix, dist_coeffs, w, h)
        self.engine.set_lut(lut)
        bl = np.ascontiguousarray(dictionary.bytesList, np.uint8)
        self.engine.set_dictionary(bl.reshape(bl.shape[0], -1), dictionary.markerSize, dictionary.maxCorrectionBits)
        self.engine.set_params(params)
        self.max_batch, self.max_markers, self.marker_length = max_batch, max_markers, float(marker_length)
        self.size = (w, h)

    @property
    def launches(self):
        return self.engine.launches

    def run_batch(self, frames, want_gray=False, want_rejected=False, marker_length=None):
        """frames: [B,H,W,3] uint8 CUDA tensor, B <= max_batch.  Returns dict of device tensors."""
        e = self.engine
        if frames.shape[0] > self.max_batch:
            raise ApseError(-1, f"batch {frames.shape[0]} exceeds max_batch {self.max_batch}")
        _, gray = e.preprocess(frames, want_bgr=False)
        det = e.detect(gray, max_markers=self.max_markers, want_rejected=want_rejected)
        ml = self.marker_length if marker_length is None else marker_length
        det["rvec"], det["tvec"] = e.pose_frames(det["corners"], det["n"], ml)
        if want_gray:
            det["gray"] = gray
        return det

    def run(self, frames, **kw):
        """Any number of frames, processed in batches of max_batch; results concatenated on the device."""
        torch = self.engine.torch
        outs = [self.run_batch(frames[i:i + self.max_batch], **kw) for i in range(0, frames.shape[0], self.max_batch)]
        return {k: torch.cat([o[k] for o in outs], 0) for k in outs[0]}

    @staticmethod
    def to_host(det):
        """Gather the small per-frame results to the host as numpy arrays (the only D2H traffic of the pipeline)."""
        out = {k: v.cpu().numpy() for k, v in det.items() if k != "gray"}
        bad = np.nonzero(out["status"])[0]
        if len(bad):
            raise ApseError(int(out["status"][bad[0]]), f"work-buffer capacity exceeded in frames {bad.tolist()[:8]}")
        return out
