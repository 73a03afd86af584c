"""GPU: the CLASSIC candidate path (north_star stages 2-4; BASELINE.json config 5) through the
C ABI against the oracle, the golden vectors of the reference's dependency, and cv2 when importable."""
import numpy as np
import pytest
from conftest import classic_cases, load_golden, classic_params, has_cv2, cv2_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from apse_uav_b200.engine import Engine
    e = Engine(0, 3840, 2160, 2)
    yield e
    e.close()


@pytest.fixture(scope="module")
def gray_dense(oracle, camera, lut, frames4k):
    K, D = camera
    mx, my = oracle.init_undistort_map(K, D, 3840, 2160)
    return oracle.preprocess(frames4k["dense"], mx, my, lut)[1]


@pytest.mark.parametrize("win,c", [(3, 7), (13, 7), (23, 7), (53, 7), (65, 7), (129, 7), (255, 7), (4, 7), (5, 3.5), (7, -2.5)])
def test_adaptive_threshold_4k_bit_exact(eng, oracle, gray_dense, win, c):
    out = eng.adaptive_threshold(gray_dense, win, c).cpu().numpy()
    assert np.array_equal(out, oracle.adaptive_threshold(gray_dense, win, c))


def test_adaptive_threshold_ragged_batch(eng, oracle):
    import torch
    rng = np.random.default_rng(5)
    for shape in ((9, 11), (33, 65), (130, 67), (361, 643)):
        g = rng.integers(0, 256, (2,) + shape, dtype=np.uint8)
        for win in (3, 23):
            out = eng.adaptive_threshold(torch.from_numpy(g).cuda(), win, 7).cpu().numpy()
            for b in range(2):
                assert np.array_equal(out[b], oracle.adaptive_threshold(g[b], win, 7)), (shape, win)


@pytest.mark.parametrize("wins", [(3, 23, 10), (13, 13, 1)])
def test_raw_quads_identical_to_oracle(eng, oracle, dictionary, gray_dense, wins):
    """candidate quads (integer corners) in the dependency's candidate order: window ascending, contour order"""
    from apse_uav_b200 import aruco
    p = classic_params(aruco, 0, wins)
    eng.set_dictionary(dictionary.raw, 4, 1)
    eng.set_params(p)
    q = eng.debug_classic(gray_dense).cpu().numpy()
    ref = oracle.classic_quads(gray_dense, p)
    assert len(ref) > 300
    assert q.shape == ref.shape and np.array_equal(q, ref)


def _detect(gray, dictionary, p):
    from apse_uav_b200 import aruco
    c, i, r = aruco.detectMarkers(gray, dictionary, parameters=p)
    return (np.array(c, np.float32).reshape(-1, 4, 2), i.ravel() if i is not None else np.zeros(0, np.int32),
            np.array(r, np.float32).reshape(-1, 4, 2))


@pytest.mark.parametrize("wins", [(3, 23, 10), (3, 53, 10), (3, 23, 4), (5, 5, 1), (13, 13, 1)])   # BASELINE.json config 5 sweep
@pytest.mark.parametrize("refine", [0, 1])
def test_detect_classic_dense_4k(oracle, dictionary, gray_dense, wins, refine):
    from apse_uav_b200 import aruco
    p = classic_params(aruco, refine, wins)
    gc, gi, gr = _detect(gray_dense, dictionary, p)
    oc, oi, orj = oracle.detect_markers_classic(gray_dense, dictionary.raw, p)
    assert len(oi) >= 150
    assert np.array_equal(gi, oi)                          # ids and order bit-exact
    assert np.array_equal(gr, orj)                         # rejected integer quads bit-exact
    if refine == 0:
        assert np.array_equal(gc, oc)                      # integer quad corners bit-exact
    else:
        assert np.abs(gc - oc).max() <= 1e-3               # sub-pixel corners within 1e-3 px (north_star)
    if has_cv2():
        import cv2
        c, i, r = cv2.aruco.ArucoDetector(cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_4X4_50), cv2_params(p)).detectMarkers(gray_dense)
        assert np.array_equal(gi, i.ravel())
        assert np.abs(gc - np.array(c, np.float32).reshape(-1, 4, 2)).max() <= (0 if refine == 0 else 1e-3)
        assert np.array_equal(gr, np.array(r, np.float32).reshape(-1, 4, 2))


@pytest.mark.parametrize("name", classic_cases())
def test_classic_golden(dictionary, name):
    """committed outputs of cv2 4.13 (tools/gen_golden_classic.py): ragged size, empty frame (574 tiny candidates -> the
    global-scratch decode path), sparse and dense frames"""
    from apse_uav_b200 import aruco
    g = load_golden(name)
    wins = tuple(int(v) for v in g["wins"])
    for refine, tag in ((0, "none"), (1, "subpix")):
        gc, gi, gr = _detect(g["gray"], dictionary, classic_params(aruco, refine, wins))
        assert np.array_equal(gi, g[f"ids_{tag}"])
        assert np.array_equal(gr, g[f"rejected_{tag}"])
        assert gc.shape == g[f"corners_{tag}"].shape
        if len(gc):
            assert np.abs(gc - g[f"corners_{tag}"]).max() <= (0 if refine == 0 else 1e-3)


def test_border_quad_swallows_group(oracle, dictionary, gray_dense):
    """same case as tests/test_oracle_classic.py: the border test comes after the grouping"""
    from apse_uav_b200 import aruco
    crop = np.ascontiguousarray(gray_dense[0:300, 800:1200])
    for rate in (0.0, 0.125):
        p = classic_params(aruco, 0, (23, 23, 1))
        p.minMarkerPerimeterRate = 38.4 / 400; p.maxMarkerPerimeterRate = 15360.5 / 400
        p.minMarkerDistanceRate = rate
        gc, gi, gr = _detect(crop, dictionary, p)
        oc, oi, orj = oracle.detect_markers_classic(crop, dictionary.raw, p)
        assert np.array_equal(gi, oi) and np.array_equal(gc, oc) and np.array_equal(gr, orj)


def test_classic_then_apriltag_same_context(oracle, dictionary, gray_dense, ref_params):
    """the two candidate paths share scratch buffers: switching modes on one context must not leak state"""
    from apse_uav_b200 import aruco
    a1 = _detect(gray_dense, dictionary, ref_params)
    _detect(gray_dense, dictionary, classic_params(aruco, 0))
    a2 = _detect(gray_dense, dictionary, ref_params)
    for x, y in zip(a1, a2):
        assert np.array_equal(x, y)
    oc, oi, orj = oracle.detect_markers_apriltag(gray_dense, dictionary.raw, ref_params)
    assert np.array_equal(a2[1], oi) and np.array_equal(a2[0], oc)


def test_pipeline_classic_mode(camera, lut, dictionary, frames4k, oracle):
    """apse_process_frames with the classic path selected (config 5 through the batched API), both frames of a batch"""
    import torch
    import apse_uav_b200 as A
    from apse_uav_b200 import aruco
    K, D = camera
    p = classic_params(aruco, 1)
    pipe = A.Pipeline(K, D, (3840, 2160), lut, dictionary, p, max_batch=2, max_markers=512)
    frames = torch.from_numpy(np.stack([frames4k["dense"], frames4k["sparse"]])).cuda()
    det = pipe.to_host(pipe.run_batch(frames, want_rejected=True))
    mx, my = oracle.init_undistort_map(K, D, 3840, 2160)
    for b, kind in enumerate(("dense", "sparse")):
        gray = oracle.preprocess(frames4k[kind], mx, my, lut)[1]
        oc, oi, orj = oracle.detect_markers_classic(gray, dictionary.raw, p)
        n = int(det["n"][b])
        assert n == len(oi) and np.array_equal(det["ids"][b, :n], oi)
        assert np.abs(det["corners"][b, :n] - oc).max() <= 1e-3
        assert int(det["n_rejected"][b]) == len(orj)
        orv, otv = oracle.estimate_pose_single_markers(oc, 0.55, K, D)
        assert np.abs(det["tvec"][b, :n] - otv[:, 0]).max() <= 1e-4 * np.abs(otv).max()
    pipe.close()


@pytest.mark.parametrize("wins", [(3, 23, 10), (13, 13, 1)])
def test_corner_refine_contour_dense_4k(oracle, dictionary, gray_dense, wins):
    """CORNER_REFINE_CONTOUR on the dense 4K frame: ids, order, rejected and the refined float32 corners
    bit-identical to the oracle (which is bit-identical to cv2 below 100 contour points per side: tests/test_oracle_classic.py);
    against cv2 itself within 0.25 px (its BLAS-backed sums above 100 points)."""
    from apse_uav_b200 import aruco
    p = classic_params(aruco, 2, wins)
    gc, gi, gr = _detect(gray_dense, dictionary, p)
    oc, oi, orj = oracle.detect_markers_classic(gray_dense, dictionary.raw, p)
    assert len(oi) >= 150
    assert np.array_equal(gi, oi) and np.array_equal(gr, orj)
    assert np.array_equal(gc, oc)
    if has_cv2():
        import cv2
        c, i, r = cv2.aruco.ArucoDetector(cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_4X4_50), cv2_params(p)).detectMarkers(gray_dense)
        cvc = np.array(c, np.float32).reshape(-1, 4, 2)
        assert np.array_equal(gi, i.ravel())
        assert np.abs(gc - cvc).max() <= 0.25
        assert (np.abs(gc - cvc).reshape(len(gc), -1).max(1) == 0).mean() >= 0.9
