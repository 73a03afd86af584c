"""GPU: the sparse evaluation of the preprocess stage (csrc/preprocess.cu, apse_preprocess_tiles_sparse /
apse_process_frames without a gray output) against the dense kernel and the oracle.

The sparse path computes the colour chain of aruco_detect.py:255-257,592 only on tiles that can matter to detectMarkers
(:267); which tiles those are is decided with a table of gray bounds per colour cell.  Checked here:
  * the bound table equals the brute-force (min, max) of the ORACLE's chain over all 2^24 colours, cell by cell;
  * every tile the detector's threshold keeps (3x3-dilated range >= aprilTagMinWhiteBlackDiff) and its one-tile ring is
    flagged, and the gray of every flagged tile is bit-identical to the dense kernel's;
  * ids, corners, rejected candidates and poses are bit-identical between the sparse and the dense path (sparse and dense
    4K frames, a smaller geometry, a low aprilTagMinWhiteBlackDiff), and equal to the oracle chain.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _engine(camera, lut, dictionary, params, w=3840, h=2160, batch=4, scale=1.0):
    from apse_uav_b200.engine import Engine
    K, D = camera
    K = K.copy()
    K[:2] *= scale
    e = Engine(0, w, h, batch)
    e.set_camera(K, D, w, h)
    e.set_lut(lut)
    bl = np.ascontiguousarray(dictionary.bytesList, np.uint8)
    e.set_dictionary(bl.reshape(bl.shape[0], -1), dictionary.markerSize, dictionary.maxCorrectionBits)
    e.set_params(params)
    return e


def test_bound_table_equals_brute_force_over_all_colours(oracle, camera, lut, dictionary, ref_params):
    e = _engine(camera, lut, dictionary, ref_params, batch=1)
    tbl = np.zeros(16 * 32 * 32, np.uint16)
    e._check(e.lib.apse_debug_sparse(e.h, tbl.ctypes.data_as(C.c_void_p), None, 0, None, e._stream()))
    idx = np.arange(1 << 24, dtype=np.uint32)
    col = np.stack([(idx >> 16) & 255, (idx >> 8) & 255, idx & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
    lab = oracle.rgb2lab(col)
    lab[..., 0] = lut[lab[..., 0]]
    gray = oracle.bgr2gray(oracle.lab2rgb(lab)).reshape(16, 16, 32, 8, 32, 8)
    lo, hi = gray.min((1, 3, 5)).reshape(-1), gray.max((1, 3, 5)).reshape(-1)
    assert np.array_equal(tbl & 255, lo)
    assert np.array_equal(255 - (tbl >> 8), hi)
    e.close()


def _dil(a, f):
    p = np.pad(a, 1, mode="edge")
    out = a.copy()
    for dy in range(3):
        for dx in range(3):
            out = f(out, p[dy:dy + a.shape[0], dx:dx + a.shape[1]])
    return out


@pytest.mark.parametrize("mwbd", [100, 40])
def test_flagged_tiles_cover_the_threshold_and_hold_exact_gray(camera, lut, dictionary, ref_params, frames4k, mwbd):
    import copy
    import torch
    p = copy.copy(ref_params)
    p.aprilTagMinWhiteBlackDiff = mwbd
    e = _engine(camera, lut, dictionary, p, batch=2)
    frames = torch.from_numpy(np.stack([frames4k["sparse"], frames4k["dense"]])).cuda()
    _, dense = e.preprocess(frames)
    gray = torch.full_like(dense, 77)
    e.preprocess_tiles(frames, gray, sparse=True)
    flags = torch.zeros((2, 540, 960), dtype=torch.uint8, device="cuda")
    n = C.c_int(0)
    e._check(e.lib.apse_debug_sparse(e.h, None, flags.data_ptr(), 2, C.byref(n), e._stream()))
    flags = flags.cpu().numpy().astype(bool)
    assert n.value == flags.sum()
    d, g = dense.cpu().numpy(), gray.cpu().numpy()
    for f in range(2):
        t = d[f].reshape(540, 4, 960, 4)
        tmin, tmax = t.min((1, 3)), t.max((1, 3))
        active = (_dil(tmax, np.maximum).astype(int) - _dil(tmin, np.minimum)) >= mwbd
        need = _dil(active, np.logical_or)
        assert not (need & ~flags[f]).any(), "a tile the detector reads was not evaluated"
        px = np.repeat(np.repeat(flags[f], 4, 0), 4, 1)
        assert np.array_equal(g[f][px], d[f][px]), "gray of an evaluated tile differs from the dense kernel"
        assert (g[f][~px] == 77).all(), "a tile outside the list was written"
    frac = flags.mean((1, 2))
    print("evaluated tile fraction (sparse frame, dense frame):", frac)
    if mwbd == 100:
        assert frac[0] < 0.15 and frac[1] < 0.5
    e.close()


def _hard_frames(w=3840, h=2160):
    """Content that stresses the bounds: saturated colour blocks, full-range colour noise, a colour ramp, black / white stripes
    one pixel wide, and a frame of uniform mid gray."""
    rng = np.random.default_rng(11)
    a = np.zeros((h, w, 3), np.uint8)
    for by in range(0, h, 120):
        for bx in range(0, w, 120):
            a[by:by + 120, bx:bx + 120] = rng.integers(0, 2, 3) * 255 if (bx // 120 + by // 120) % 3 else rng.integers(0, 256, 3)
    b = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    c = np.zeros((h, w, 3), np.uint8)
    c[..., 0] = (np.arange(w) * 255 // (w - 1))[None, :]
    c[..., 1] = (np.arange(h) * 255 // (h - 1))[:, None]
    c[..., 2] = ((np.arange(w)[None, :] + np.arange(h)[:, None]) % 512 // 2).astype(np.uint8)
    c[:, ::7] = 255
    c[::5, :] = 0
    d = np.full((h, w, 3), 128, np.uint8)
    d += rng.integers(0, 4, (h, w, 3), dtype=np.uint8)
    return np.stack([a, b, c, d])


def test_tile_bounds_contain_the_dense_gray(camera, lut, dictionary, ref_params, frames4k):
    """The bounds pass (k_preprocess_tma MODE 1: every pixel remapped, its colour cell looked up in the (min, max) table of the
    chain) is rigorous: for EVERY tile of every frame [lo, hi] contains the gray the dense kernel computes
    (aruco_detect.py:250-259,592) -- markers, saturated colours, full-range noise, one-pixel stripes, the black margin the
    undistortion leaves at the frame border."""
    import torch
    e = _engine(camera, lut, dictionary, ref_params, batch=3)
    hard = _hard_frames()
    for frames in (np.stack([frames4k["sparse"], frames4k["dense"], hard[0]]), hard[1:]):
        ft = torch.from_numpy(frames).cuda()
        _, dense = e.preprocess(ft)
        gray = torch.empty_like(dense)
        e.preprocess_tiles(ft, gray, sparse=True)
        tb = torch.zeros((len(frames), 540, 960), dtype=torch.int16, device="cuda")
        e._check(e.lib.apse_debug_tile_bounds(e.h, tb.data_ptr(), len(frames), e._stream()))
        torch.cuda.synchronize()
        tb = tb.cpu().numpy().view(np.uint16)
        lo, hi = (tb & 255).astype(int), (tb >> 8).astype(int)
        t = dense.cpu().numpy().reshape(len(frames), 540, 4, 960, 4)
        tmin, tmax = t.min((2, 4)).astype(int), t.max((2, 4)).astype(int)
        assert (lo <= tmin).all(), f"lower bound above the exact minimum in {(lo > tmin).sum()} tiles"
        assert (hi >= tmax).all(), f"upper bound below the exact maximum in {(hi < tmax).sum()} tiles"
        print("mean bound slack (lo, hi) per frame:", (tmin - lo).mean((1, 2)), (hi - tmax).mean((1, 2)))
    e.close()


def _both(e, frames, max_markers=256):
    import torch
    outs = []
    for gray in (None, torch.empty(frames.shape[:3], dtype=torch.uint8, device="cuda")):
        det = e.alloc_detections(frames.shape[0], max_markers, True, pose=True)
        e.process_frames(frames, det, 0.55, gray=gray)     # gray=None: sparse evaluation; with a gray output: dense kernel
        torch.cuda.synchronize()
        outs.append({k: v.cpu().numpy() for k, v in det.items()})
    return outs


@pytest.mark.parametrize("mwbd", [100, 70])
def test_sparse_detections_bit_identical_to_dense(camera, lut, dictionary, ref_params, frames4k, oracle, mwbd):
    import copy
    import torch
    p = copy.copy(ref_params)
    p.aprilTagMinWhiteBlackDiff = mwbd
    e = _engine(camera, lut, dictionary, p, batch=3)
    from tools import synth
    third = synth.make_frame(dictionary.bytesList, 91, ids=(1, 2, 3, 4, 7, 9, 11), side_range=(30, 120), noise_sigma=4.0, occlude_frac=0.3)
    frames = torch.from_numpy(np.stack([frames4k["sparse"], frames4k["dense"], third])).cuda()
    sparse, dense = _both(e, frames)
    assert (dense["status"] == 0).all() and (sparse["status"] == 0).all()
    assert dense["n"][0] >= 4 and dense["n"][1] >= 100
    for k in ("n", "ids", "corners", "n_rejected", "rejected", "rvec", "tvec"):
        assert np.array_equal(sparse[k], dense[k]), k
    if mwbd == 100:   # and the oracle chain agrees (sparse frame)
        K, D = camera
        ox, oy = oracle.init_undistort_map(K, D, 3840, 2160)
        _, g = oracle.preprocess(frames4k["sparse"], ox, oy, lut)
        oc, oi, _ = oracle.detect_markers_apriltag(g, dictionary.raw, p)
        n = int(sparse["n"][0])
        assert n == len(oi) and np.array_equal(sparse["ids"][0, :n], oi) and np.abs(sparse["corners"][0, :n] - oc).max() <= 1e-3
    e.close()


def test_sparse_small_geometry_and_fallback(camera, lut, dictionary, ref_params):
    """1920x1080 (TMA path, run-time width) and 1000x720 (no TMA path: the sparse entry falls back to the dense kernel)."""
    import torch
    from tools import synth
    for (w, h), scale in (((1920, 1080), 0.5), ((1000, 720), 0.26)):
        e = _engine(camera, lut, dictionary, ref_params, w=w, h=h, batch=2, scale=scale)
        frames = np.stack([synth.make_frame(dictionary.bytesList, 5 + i, w, h, ids=(1, 2, 3, 4, 7), side_range=(40, 80)) for i in range(2)])
        sparse, dense = _both(e, torch.from_numpy(frames).cuda(), 64)
        assert dense["n"].min() >= 3
        for k in ("n", "ids", "corners", "n_rejected", "rejected", "rvec", "tvec"):
            assert np.array_equal(sparse[k], dense[k]), (w, h, k)
        e.close()


def test_stale_sparse_gray_is_refused(camera, lut, dictionary, ref_params, frames4k):
    """The gray buffer of a sparse batch is only complete together with its tile flags: a detect call after the flags were
    dropped (set_params in between) fails loudly instead of reading pixels that were never computed."""
    import torch
    from apse_uav_b200._lib import ApseError
    e = _engine(camera, lut, dictionary, ref_params, batch=1)
    frames = torch.from_numpy(frames4k["sparse"][None]).cuda()
    gray = torch.empty((1, 2160, 3840), dtype=torch.uint8, device="cuda")
    det = e.alloc_detections(1, 64, True, pose=True)
    e.preprocess_tiles(frames, gray, sparse=True)
    e.set_params(ref_params)
    with pytest.raises(ApseError):
        e.detect_pose_frames(gray, det, 0.55)
    e.preprocess_tiles(frames, gray, sparse=True)
    e.detect_pose_frames(gray, det, 0.55)
    torch.cuda.synchronize()
    assert int(det["n"][0]) >= 4
    e.close()
