import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_cv2():
    try:
        import cv2  # noqa: F401
        return True
    except ImportError:
        return False


needs_cv2 = pytest.mark.skipif(not has_cv2(), reason="cv2 (the reference's dependency) not importable")


@pytest.fixture(scope="session")
def camera():
    cam = json.load(open(os.path.join(GOLDEN, "cam_params.json")))
    return np.array(cam["mtx"]), np.array(cam["dist"]).ravel()


@pytest.fixture(scope="session")
def lut():
    import __graft_entry__ as G
    return G.gamma_lut()


@pytest.fixture(scope="session")
def dictionary():
    from apse_uav_b200 import aruco
    return aruco.getPredefinedDictionary(aruco.DICT_4X4_50)


@pytest.fixture(scope="session")
def ref_params():
    from apse_uav_b200 import aruco
    import __graft_entry__ as G
    return G.reference_parameters(aruco)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


def golden_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith("classic_"))


def classic_cases():
    """fixtures of the classic candidate path (tools/gen_golden_classic.py)"""
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f.startswith("classic_"))


def classic_params(aruco_mod, refine, wins=(3, 23, 10)):
    """parameters of aruco_detect.py:190-203 with the classic candidate path selected (NONE = 0 / SUBPIX = 1)"""
    import __graft_entry__ as G
    p = G.reference_parameters(aruco_mod)
    p.cornerRefinementMethod = refine
    p.adaptiveThreshWinSizeMin, p.adaptiveThreshWinSizeMax, p.adaptiveThreshWinSizeStep = wins
    return p


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def frames4k(dictionary):
    """One sparse and one dense synthetic 4K frame (seeded)."""
    from tools import synth
    return {"sparse": synth.make_frame(dictionary.bytesList, 3), "dense": synth.make_dense_frame(dictionary.bytesList, 11)}


def cv2_params(p):
    import cv2
    q = cv2.aruco.DetectorParameters()
    for k, v in vars(p).items():
        setattr(q, k, v)
    return q


# ---- CSV rows of aruco_detect.py:146-185: comparison with the reference script's own output -----------------------------
CSV_EXACT_COLS = [0, 1, 3, 7, 10, 13]                       # frame id, detection flags, leds_ID: integers, must be equal
# one unit in the last place the reference prints (round(.., 5 / 2 / 3)) on top of north_star's 1e-4 relative bar
CSV_QUANTUM = {2: 1e-5, 4: 1e-2, 5: 1e-2, 6: 1e-2, 8: 1e-3, 9: 1e-3, 11: 1e-3, 12: 1e-3, 14: 1e-3, 15: 1e-3}


def golden_csv(name):
    g = json.load(open(os.path.join(GOLDEN, name)))
    return g, np.array([[float(v) for v in line.split(",")] for line in g["csv"][1:]])


def golden_events(g):
    ev = g.get("events")
    if not ev:
        return None
    return {int(k): {"hide": v["hide"], "jump": {int(i): tuple(x) for i, x in v["jump"].items()}} for k, v in ev.items()}


def assert_csv_rows_match(rows, ref):
    """rows: list of row dicts (postpass.CSV_FIELDS) or an (n,16) array; ref: the reference script's rows as floats."""
    from apse_uav_b200.postpass import CSV_FIELDS
    got = rows if isinstance(rows, np.ndarray) else np.array([[float(r[f]) for f in CSV_FIELDS] for r in rows])
    assert got.shape == ref.shape
    assert np.array_equal(got[:, CSV_EXACT_COLS], ref[:, CSV_EXACT_COLS]), "frame ids / detection flags / leds_ID differ"
    for col, q in CSV_QUANTUM.items():
        err = np.abs(got[:, col] - ref[:, col])
        tol = 1e-4 * np.abs(ref[:, col]) + q * 1.0001
        assert np.all(err <= tol), (col, int(np.argmax(err - tol)), got[:, col][np.argmax(err - tol)], ref[:, col][np.argmax(err - tol)])
