"""GPU: annotated frames (aruco_detect.py:421-425,494-500,614-616).  The overlay rasteriser against cv2.line /
cv2.circle on the same primitives, and the overlays of a real sequence (marker quads, vehicle outlines, distance lines) against
the same primitives drawn by cv2: IoU of the touched area >= 0.93, identical colours on at least 97 % of the pixels both touched
(cv2 fills thick lines through its 16.16 fixed-point polygon filler; single edge pixels differ)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def draw_cv2(frames_host, prims):
    """The same primitives through cv2 on host frames (test reference for the rasteriser)."""
    import cv2
    for p in prims[np.argsort(prims["frame"], kind="stable")]:
        img = frames_host[int(p["frame"])]
        col = tuple(int(c) for c in p["bgr"][:3])
        if p["kind"] == 0:
            cv2.line(img, (int(p["x0"]), int(p["y0"])), (int(p["x1"]), int(p["y1"])), col, int(p["thickness"]))
        else:
            cv2.circle(img, (int(p["x0"]), int(p["y0"])), int(p["thickness"]), col, -1)
    return frames_host


def _compare(got, want, base):
    touched_g, touched_w = (got != base).any(-1), (want != base).any(-1)
    inter, union = (touched_g & touched_w).sum(), (touched_g | touched_w).sum()
    assert union > 0 and inter / union >= 0.93, inter / union
    same = (got == want).all(-1)
    assert same[touched_g & touched_w].mean() >= 0.97     # where both drew, the same primitive is on top
    return inter / union


def test_rasteriser_matches_cv2_on_random_primitives(camera, lut, dictionary, ref_params):
    pytest.importorskip("cv2")
    import torch
    from apse_uav_b200 import render
    from apse_uav_b200._lib import OVERLAY_PRIM_DTYPE
    from apse_uav_b200.engine import default_engine
    e = default_engine(1280, 720)
    rng = np.random.default_rng(5)
    prims = []
    for f in range(3):
        for _ in range(40):
            x0, y0, x1, y1 = rng.integers(-50, 1330), rng.integers(-50, 770), rng.integers(-50, 1330), rng.integers(-50, 770)
            prims.append((f, 0, x0, y0, x1, y1, int(rng.choice([3, 4, 5, 9])), tuple(int(v) for v in rng.integers(1, 255, 3)) + (0,)))
        for _ in range(10):
            prims.append((f, 1, rng.integers(0, 1280), rng.integers(0, 720), 0, 0, int(rng.integers(1, 12)), tuple(int(v) for v in rng.integers(1, 255, 3)) + (0,)))
    prims = np.array(prims, dtype=OVERLAY_PRIM_DTYPE)
    base = np.zeros((3, 720, 1280, 3), np.uint8)
    got = render.draw(e, torch.from_numpy(base.copy()).cuda(), prims).cpu().numpy()
    want = draw_cv2(base.copy(), prims)
    for f in range(3):
        _compare(got[f], want[f], base[f])


def test_sequence_overlays(camera, lut, dictionary, ref_params):
    pytest.importorskip("cv2")
    import torch
    import apse_uav_b200 as A
    from apse_uav_b200 import render, sequence
    from tools import synth
    K, D = camera
    pipe = A.Pipeline(K, D, (3840, 2160), lut, dictionary, ref_params, max_batch=4, max_markers=64)
    frames_h = np.stack(list(synth.make_sequence(dictionary.bytesList, 500, 4)))
    frames = torch.from_numpy(frames_h).cuda()
    det = pipe.run_sequence(frames)
    info = sequence.postpass_device(pipe.engine, det, details=True)
    prims = render.sequence_primitives(info)
    per_frame = np.bincount(prims["frame"], minlength=4)
    assert (per_frame >= 4 * 4 + 3 * (4 + 2 + 1)).all()      # 4 marker quads, 3 x (outline, two lines, point) in every frame
    out = render.annotate_sequence(pipe.engine, frames.clone(), info).cpu().numpy()
    want = draw_cv2(frames_h.copy(), prims)
    for f in range(4):
        _compare(out[f], want[f], frames_h[f])
    # the red line of every distance job ends at the vehicle marker's centre, the yellow one on the vehicle's outline
    for row in info["rows"]:
        for v in range(3):
            j = int(row["job_dist"][v])
            assert j >= 0 and info["results"][j]["valid"] == 1
            o = info["results"][j]["outline_px"]
            p = info["results"][j]["nearest_px"]
            assert o[:, 0].min() - 1 <= p[0] <= o[:, 0].max() + 1 and o[:, 1].min() - 1 <= p[1] <= o[:, 1].max() + 1
    pipe.close()
