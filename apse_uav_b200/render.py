"""Annotated frames of aruco_detect.py on the GPU: the overlays the reference draws with cv2 before it shows
or saves a frame -- marker quads (drawMarkers, :614-616), vehicle outlines (drawBoundingBox, :421-425), the two distance lines
(drawLinesOnImage, :494-500) and the projected points (:378-380) -- rasterised into device-resident BGR frames by
libapse_b200 (apse_draw_overlay, csrc/render.cu), from the results of the native sequence post-pass.  Text (cv2.putText) is
left to the host: pass the annotated frame to the reference's own printDataOnImage if it is wanted.

    det  = pipe.run_sequence(frames)
    info = sequence.postpass_device(pipe.engine, det, details=True)
    render.annotate_sequence(pipe.engine, frames, info)          # frames now carry the overlays
"""
from __future__ import annotations

import numpy as np

from ._lib import ApseError, OVERLAY_PRIM_DTYPE

GREEN, BLUE, YELLOW, RED, CYAN = (0, 255, 0), (255, 0, 0), (0, 255, 255), (0, 0, 255), (255, 255, 0)


def _seg(out, frame, p, q, thickness, colour):
    out.append((frame, 0, int(p[0]), int(p[1]), int(q[0]), int(q[1]), thickness, colour + (0,)))


def _closed(out, frame, pts, thickness, colour):
    for a in range(len(pts)):
        _seg(out, frame, pts[a], pts[(a + 1) % len(pts)], thickness, colour)


def sequence_primitives(info, draw_markers=True, draw_outlines=True, draw_lines=True, draw_points=True):
    """Overlay primitives of a whole sequence, in the reference's drawing order per frame."""
    rows, jobs, results = info["rows"], info["jobs"], info["results"]
    prims = []
    for fr, row in enumerate(rows):
        if draw_markers:   # cv2.drawContours(frame, [np.maximum(0, np.int32(corners))], -1, (0,255,0), 3)
            for i in range(int(info["n"][fr])):
                if (int(row["accepted_mask"]) >> i) & 1:
                    _closed(prims, fr, np.maximum(0, info["corners"][fr, i].astype(np.int32)), 3, GREEN)
        for v in range(3):
            j = int(row["job_dist"][v])
            if j < 0 or not results[j]["valid"]:
                continue
            J, R = jobs[j], results[j]
            if draw_outlines:   # cv2.drawContours(frame, [imgpts[0:4]], -1, (255,0,0), 5)
                _closed(prims, fr, R["outline_px"], 5, BLUE)
            src = (int(J["src"][0]), int(J["src"][1]))
            if draw_lines:      # drawLinesOnImage: yellow to the closest outline point, red to the vehicle's marker
                _seg(prims, fr, src, R["nearest_px"], 5, YELLOW)
                _seg(prims, fr, src, (int(J["tgt"][0]), int(J["tgt"][1])), 5, RED)
            if draw_points:     # cv2.circle(frame, point, 5, (255,255,0), -1)
                prims.append((fr, 1, int(R["nearest_px"][0]), int(R["nearest_px"][1]), 0, 0, 5, CYAN + (0,)))
    return np.array(prims, dtype=OVERLAY_PRIM_DTYPE) if prims else np.zeros(0, OVERLAY_PRIM_DTYPE)


def draw(engine, frames, prims):
    """apse_draw_overlay: prims (numpy OVERLAY_PRIM_DTYPE, any order) into frames [B,H,W,3] uint8 CUDA tensor, in place."""
    torch = engine.torch
    if frames.dim() != 4 or frames.shape[3] != 3 or frames.dtype != torch.uint8 or not frames.is_contiguous() or not frames.is_cuda:
        raise ApseError(-1, "draw: frames must be a contiguous uint8 CUDA tensor [B,H,W,3]")
    if len(prims) == 0:
        return frames
    prims = np.ascontiguousarray(prims[np.argsort(prims["frame"], kind="stable")])
    if prims["frame"].min() < 0 or prims["frame"].max() >= frames.shape[0]:
        raise ApseError(-1, "draw: primitive outside the batch")
    d = torch.from_numpy(prims.view(np.uint8).reshape(-1).copy()).to(engine.tdev)
    B, H, W, _ = frames.shape
    rc = engine.lib.apse_draw_overlay(engine.h, frames.data_ptr(), W, H, B, d.data_ptr(), len(prims), engine._stream())
    if rc != 0:
        raise ApseError(rc, engine.lib.apse_last_error(engine.h).decode())
    return frames


def annotate_sequence(engine, frames, info, **kw):
    return draw(engine, frames, sequence_primitives(info, **kw))
