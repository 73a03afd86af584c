"""GPU parity: aruco.detectMarkers in CORNER_REFINE_APRILTAG mode (aruco_detect.py:266-267) through the C ABI.
Bars (BASELINE.json): ids / candidate counts bit-exact, sub-pixel corners <= 1e-3 px (observed: bit-identical)."""
import numpy as np
import pytest
from conftest import golden_cases, load_golden, has_cv2, cv2_params

pytestmark = pytest.mark.gpu
CORNER_TOL = 1e-3


@pytest.fixture(scope="module")
def eng(dictionary, ref_params):
    from apse_uav_b200.engine import Engine
    e = Engine(0, 3840, 2160, 4)
    e.set_dictionary(dictionary.raw, dictionary.markerSize, dictionary.maxCorrectionBits)
    e.set_params(ref_params)
    yield e
    e.close()


@pytest.fixture(scope="module")
def grays4k(oracle, camera, lut, frames4k):
    K, D = camera
    ox, oy = oracle.init_undistort_map(K, D, 3840, 2160)
    return {k: oracle.preprocess(v, ox, oy, lut)[1] for k, v in frames4k.items()}


@pytest.mark.parametrize("kind", ["sparse", "dense"])
def test_apriltag_stages_4k(eng, oracle, ref_params, grays4k, kind):
    """every intermediate of the candidate path: ternary image, component labels, point / cluster counts, raw quads."""
    import torch
    gray = grays4k[kind]
    oq, st = oracle.at_quads(gray, ref_params, dumps=True)
    dbg = eng.debug_apriltag(torch.from_numpy(gray).cuda())
    th = dbg["thresh"].cpu().numpy()
    assert np.array_equal(th, st["thresh"])
    lab = dbg["labels"].cpu().numpy().view(np.uint32)
    m = th != 127
    assert np.array_equal(lab[m], st["rep"][m])          # same representative (smallest pixel index) everywhere
    assert (dbg["points"], dbg["clusters"], dbg["fitted"], dbg["n_quads"]) == (st["points"], st["clusters"], st["fitted"], len(oq))
    gq = dbg["quads"].cpu().numpy().reshape(-1, 8)
    dist = np.abs(gq[:, None, :] - oq.reshape(-1, 8)[None, :, :]).max(-1)
    assert (dist.min(1) == 0).all() and (dist.min(0) == 0).all()   # bit-identical quad set


@pytest.mark.parametrize("kind", ["sparse", "dense"])
def test_detect_markers_4k(oracle, dictionary, ref_params, grays4k, kind):
    from apse_uav_b200 import aruco
    gray = grays4k[kind]
    c, ids, rej = aruco.detectMarkers(gray, dictionary, parameters=ref_params)
    oc, oi, orj = oracle.detect_markers_apriltag(gray, dictionary.raw, ref_params)
    assert ids.dtype == np.int32 and ids.shape == (len(oi), 1) and c[0].shape == (1, 4, 2) and c[0].dtype == np.float32
    assert np.array_equal(ids.ravel(), oi)
    assert np.abs(np.array(c).reshape(-1, 4, 2) - oc).max() <= CORNER_TOL
    assert len(rej) == len(orj) and (len(orj) == 0 or np.abs(np.array(rej).reshape(-1, 4, 2) - orj).max() <= CORNER_TOL)
    if has_cv2():
        import cv2
        det = cv2.aruco.ArucoDetector(cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_4X4_50), cv2_params(ref_params))
        rc, ri, rr = det.detectMarkers(gray)
        assert np.array_equal(ids.ravel(), ri.ravel()) and len(rr) == len(rej)
        assert np.abs(np.array(c) - np.array(rc)).max() <= CORNER_TOL


def test_empty_frame_returns_none_ids(dictionary, ref_params):
    from apse_uav_b200 import aruco
    flat = np.full((480, 640), 128, np.uint8)
    c, ids, rej = aruco.detectMarkers(flat, dictionary, parameters=ref_params)
    assert ids is None and c == () and rej == ()    # aruco_detect.py:599 relies on ids being None
    noise = np.random.default_rng(0).integers(100, 140, (480, 640), dtype=np.uint8)
    c, ids, rej = aruco.detectMarkers(noise, dictionary, parameters=ref_params)
    assert ids is None


def test_ragged_sizes_borders_and_bgr_input(oracle, dictionary, ref_params):
    """sizes not divisible by the 4x4 threshold tiles / the CCL tiles, markers cut by the image border, BGR input."""
    from apse_uav_b200 import aruco
    from tools import synth
    for seed, (w, h) in enumerate([(1280, 720), (1283, 721), (1001, 750), (333, 251)]):
        f = synth.make_frame(dictionary.bytesList, 40 + seed, w, h, ids=[i % 50 for i in range(30)], side_range=(40, 120),
                             jitter=0.15, occlude_frac=0.15, margin=0, noise_sigma=3)
        gray = oracle.bgr2gray(f)
        oc, oi, orj = oracle.detect_markers_apriltag(gray, dictionary.raw, ref_params)
        for img in (gray, f):
            c, ids, rej = aruco.detectMarkers(img, dictionary, parameters=ref_params)
            got = ids.ravel() if ids is not None else np.zeros(0, np.int32)
            assert np.array_equal(got, oi), (w, h)
            assert len(oi) == 0 or np.abs(np.array(c).reshape(-1, 4, 2) - oc).max() <= CORNER_TOL
            assert len(rej) == len(orj)


def test_batch_equals_single_frames(eng, grays4k):
    import torch
    batch = torch.from_numpy(np.stack([grays4k["sparse"], grays4k["dense"], grays4k["sparse"], grays4k["dense"]])).cuda()
    r = eng.detect(batch, max_markers=256)
    n = r["n"].cpu().numpy()
    assert (r["status"].cpu().numpy() == 0).all() and n[0] == n[2] and n[1] == n[3] and n[0] >= 4 and n[1] >= 150
    for a, b in ((0, 2), (1, 3)):
        assert torch.equal(r["ids"][a], r["ids"][b]) and torch.equal(r["corners"][a], r["corners"][b])
    single = eng.detect(batch[1:2], max_markers=256)
    assert torch.equal(single["ids"][0], r["ids"][1]) and torch.equal(single["corners"][0], r["corners"][1])


def test_capacity_overflow_is_reported_not_truncated(eng, grays4k):
    import torch
    r = eng.detect(torch.from_numpy(grays4k["dense"]).cuda(), max_markers=16)
    assert int(r["status"][0]) == -4 and int(r["n"][0]) == 16


def test_parameter_validation_mirrors_opencv_asserts(dictionary):
    from apse_uav_b200 import aruco, ApseError
    img = np.zeros((64, 64), np.uint8)
    p = aruco.DetectorParameters(); p.cornerRefinementMethod = aruco.CORNER_REFINE_APRILTAG; p.markerBorderBits = 0
    with pytest.raises(ApseError):
        aruco.detectMarkers(img, dictionary, parameters=p)
    p = aruco.DetectorParameters(); p.cornerRefinementMethod = aruco.CORNER_REFINE_APRILTAG; p.aprilTagQuadDecimate = 16.0   # 64 / 16: nothing left to look at
    with pytest.raises(ApseError):
        aruco.detectMarkers(img, dictionary, parameters=p)
    p = aruco.DetectorParameters(); p.cornerRefinementMethod = aruco.CORNER_REFINE_APRILTAG; p.aprilTagDeglitch = 1
    with pytest.raises(ApseError):
        aruco.detectMarkers(img, dictionary, parameters=p)
    with pytest.raises(ApseError):
        aruco.detectMarkers(np.zeros((0, 0), np.uint8), dictionary)
    with pytest.raises(ApseError):
        aruco.detectMarkers(img.astype(np.float32), dictionary)


@pytest.mark.parametrize("name", golden_cases())
def test_detect_golden(dictionary, ref_params, oracle, name):
    from apse_uav_b200 import aruco
    g = load_golden(name)
    h, w = g["frame"].shape[:2]
    ox, oy = oracle.init_undistort_map(g["K"], g["D"], w, h)
    _, gray = oracle.preprocess(g["frame"], ox, oy, g["lut"])
    c, ids, rej = aruco.detectMarkers(gray, dictionary, parameters=ref_params)
    got = ids.ravel() if ids is not None else np.zeros(0, np.int32)
    assert np.array_equal(got, g["ids"])
    assert len(got) == 0 or np.abs(np.array(c).reshape(-1, 4, 2) - g["corners"]).max() <= CORNER_TOL
    assert len(rej) == len(g["rejected"])


@pytest.mark.parametrize("dec,sigma", [(1.5, 0.0), (2.0, 0.0), (3.0, 0.0), (0.0, 0.8), (0.0, -0.8), (2.0, 0.8), (4.0, -1.3), (7.0, 0.0), (2.5, 1.3)])
def test_quad_decimate_and_sigma(oracle, camera, lut, dictionary, ref_params, frames4k, dec, sigma):
    """aprilTagQuadDecimate / aprilTagQuadSigma (aruco_detect.py:203,231-233) on 4K frames: ids, order, float32
    corners and rejected count equal to the oracle (which equals cv2: tests/test_oracle_detect.py), and to cv2 itself."""
    import copy
    import torch
    from apse_uav_b200 import aruco
    from apse_uav_b200.engine import Engine
    p = copy.copy(ref_params)
    p.aprilTagQuadDecimate, p.aprilTagQuadSigma = dec, sigma
    e = Engine(0, 3840, 2160, 2)
    bl = np.ascontiguousarray(dictionary.bytesList, np.uint8)
    e.set_dictionary(bl.reshape(bl.shape[0], -1), dictionary.markerSize, dictionary.maxCorrectionBits)
    e.set_params(p)
    K, D = camera
    e.set_camera(K, D, 3840, 2160)
    e.set_lut(lut)
    _, g = e.preprocess(torch.from_numpy(np.stack([frames4k["sparse"], frames4k["dense"]])).cuda())   # aruco_detect.py:250-259,592
    grays = g.cpu().numpy()
    det = e.detect(g, max_markers=512)
    res = {k: v.cpu().numpy() for k, v in det.items()}
    assert (res["status"] == 0).all()
    for f in range(2):
        oc, oi, orj = oracle.detect_markers_apriltag(grays[f], dictionary.raw, p)
        n = int(res["n"][f])
        assert n == len(oi) and n >= (4 if f == 0 else 60) or dec >= 7 and n == len(oi)
        assert np.array_equal(res["ids"][f, :n], oi)
        assert np.array_equal(res["corners"][f, :n], oc)
        assert int(res["n_rejected"][f]) == len(orj)
    try:
        import cv2
        from conftest import cv2_params
    except ImportError:
        return
    det2 = cv2.aruco.ArucoDetector(cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_4X4_50), cv2_params(p))
    cc, ci, _ = det2.detectMarkers(grays[0])
    n = int(res["n"][0])
    assert ci.ravel().tolist() == res["ids"][0, :n].tolist()
    assert np.abs(np.array([c[0] for c in cc]) - res["corners"][0, :n]).max() <= 1e-3
    e.close()


@pytest.mark.parametrize("mode", [3, 0])
def test_nested_markers(oracle, dictionary, ref_params, mode):
    """Markers inside markers: the quad that encloses an already decoded marker is identified as cv2 4.13 does (ids, order,
    corners, rejected count equal to the oracle = cv2, tests/test_oracle_detect.py::test_nested_markers_vs_cv2)."""
    import copy
    import torch
    from apse_uav_b200.engine import Engine
    from tools import synth
    p = copy.copy(ref_params)
    p.cornerRefinementMethod = mode
    e = Engine(0, 1920, 1080, 6)
    bl = np.ascontiguousarray(dictionary.bytesList, np.uint8)
    e.set_dictionary(bl.reshape(bl.shape[0], -1), dictionary.markerSize, dictionary.maxCorrectionBits)
    e.set_params(p)
    frames = [synth.make_nested_frame(dictionary.bytesList, 100 + s, levels=2 + s % 2)[0] for s in (1, 2, 5, 9, 10, 11)]
    grays = np.stack([np.ascontiguousarray(((f[..., 0].astype(np.int32) * 3735 + f[..., 1].astype(np.int32) * 19235 + f[..., 2].astype(np.int32) * 9798 + 16384) >> 15).astype(np.uint8)) for f in frames])
    det = e.detect(torch.from_numpy(grays).cuda(), max_markers=128)
    res = {k: v.cpu().numpy() for k, v in det.items()}
    assert (res["status"] == 0).all()
    fn = oracle.detect_markers_apriltag if mode == 3 else oracle.detect_markers_classic
    for f in range(len(frames)):
        oc, oi, orj = fn(grays[f], dictionary.raw, p)
        n = int(res["n"][f])
        assert n == len(oi) and np.array_equal(res["ids"][f, :n], oi) and np.array_equal(res["corners"][f, :n], oc)
        assert int(res["n_rejected"][f]) == len(orj)
    e.close()


def test_identification_stage_in_isolation(oracle, camera, lut, dictionary, ref_params, frames4k):
    """a6.A6 (_extractBits) and a6.A7 (Dictionary::identify) on their own (apse_debug_decode): for every raw quad of the dense 4K
    frame and for random / degenerate quads, the 48 x 48 canonical image (nearest-neighbour warp), the Otsu threshold, the 36 cell
    bits and (valid, id, rotation) equal the CPU restatement -- a decode error that cancels end to end would show up here."""
    import ctypes as C
    import torch
    from apse_uav_b200.engine import Engine
    K, D = camera
    e = Engine(0, 3840, 2160, 1)
    bl = np.ascontiguousarray(dictionary.bytesList, np.uint8)
    e.set_dictionary(bl.reshape(bl.shape[0], -1), dictionary.markerSize, dictionary.maxCorrectionBits)
    e.set_params(ref_params)
    e.set_camera(K, D, 3840, 2160)
    e.set_lut(lut)
    _, g = e.preprocess(torch.from_numpy(frames4k["dense"][None]).cuda())
    gray = g[0].cpu().numpy()
    quads = oracle.at_quads(gray, ref_params)
    rng = np.random.default_rng(9)
    extra = []
    for _ in range(40):   # random convex-ish quads anywhere, partly outside the frame, some tiny
        c = rng.uniform([-40, -40], [3880, 2200]); s = rng.uniform(3, 200); a = rng.uniform(0, 6.28)
        base = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]]) * s / 2 + rng.uniform(-0.2 * s, 0.2 * s, (4, 2))
        extra.append(base @ np.array([[np.cos(a), np.sin(a)], [-np.sin(a), np.cos(a)]]) + c)
    quads = np.concatenate([quads, np.float32(extra)]).astype(np.float32)
    n = len(quads)
    assert n >= 150
    dq = torch.from_numpy(quads.reshape(n, 8).copy()).cuda()
    img = torch.zeros((n, 64 * 64), dtype=torch.uint8, device="cuda")
    bits = torch.zeros((n, 256), dtype=torch.uint8, device="cuda")
    res = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
    e._check(e.lib.apse_debug_decode(e.h, g.data_ptr(), 3840, 2160, dq.data_ptr(), n, img.data_ptr(), bits.data_ptr(), res.data_ptr(), e._stream()))
    img, bits, res = img.cpu().numpy(), bits.cpu().numpy(), res.cpu().numpy()
    dp = oracle.DecParams.from_cv(ref_params)
    nb, S = 6, 48
    n_valid = n_otsu = 0
    for i in range(n):
        assert np.array_equal(img[i, :S * S].reshape(S, S), oracle.warp_nearest(gray, quads[i], S)), i
        obits, othr = oracle.extract_bits(gray, quads[i], dp)
        assert int(res[i, 3]) == othr, (i, res[i, 3], othr)
        assert np.array_equal(bits[i, :nb * nb].reshape(nb, nb), obits), i
        n_otsu += othr >= 0
        # identification = the end-to-end oracle on this single candidate
        oc, oi, _ = oracle.identify_candidates(gray, quads[i:i + 1], oracle.DecParams.from_cv(ref_params, ), dictionary.raw)
        border_ok = bool(np.all((quads[i] >= 3) & (quads[i] < [3840 - 3, 2160 - 3])))   # minDistanceToBorder is applied later, not by the tap
        if border_ok:
            assert bool(res[i, 0]) == (len(oi) == 1), i
            if len(oi) == 1:
                assert int(res[i, 1]) == int(oi[0])
                assert np.array_equal(np.roll(quads[i], int(res[i, 2]), axis=0), oc[0])
                n_valid += 1
    assert n_valid >= 100 and n_otsu >= 150
    e.close()


@pytest.mark.parametrize("mode", [3, 0])
def test_inverted_markers(oracle, dictionary, ref_params, mode):
    """detectInvertedMarker on the GPU: ids, order, corners and rejected count equal to the oracle (= cv2 4.13,
    tests/test_oracle_detect.py::test_inverted_markers_vs_cv2) on frames where every other marker is white on black -- the
    inverted-bits branch of the identification and the descending member order of the too-close groups."""
    import copy
    import torch
    from apse_uav_b200.engine import Engine
    from tools import synth
    p = copy.copy(ref_params)
    p.cornerRefinementMethod = mode
    p.detectInvertedMarker = True
    e = Engine(0, 1920, 1080, 6)
    bl = np.ascontiguousarray(dictionary.bytesList, np.uint8)
    e.set_dictionary(bl.reshape(bl.shape[0], -1), dictionary.markerSize, dictionary.maxCorrectionBits)
    e.set_params(p)
    frames = [synth.make_inverted_frame(dictionary.bytesList, 500 + s, n_markers=40 if s % 2 else 6) for s in range(6)]
    grays = np.stack([np.ascontiguousarray(((f[..., 0].astype(np.int32) * 3735 + f[..., 1].astype(np.int32) * 19235 + f[..., 2].astype(np.int32) * 9798 + 16384) >> 15).astype(np.uint8)) for f in frames])
    det = e.detect(torch.from_numpy(grays).cuda(), max_markers=256)
    res = {k: v.cpu().numpy() for k, v in det.items()}
    assert (res["status"] == 0).all()
    fn = oracle.detect_markers_apriltag if mode == 3 else oracle.detect_markers_classic
    total = 0
    for f in range(len(frames)):
        oc, oi, orj = fn(grays[f], dictionary.raw, p)
        n = int(res["n"][f])
        assert n == len(oi) and np.array_equal(res["ids"][f, :n], oi) and np.array_equal(res["corners"][f, :n], oc)
        assert int(res["n_rejected"][f]) == len(orj)
        total += n
    assert total > 50
    e.close()
