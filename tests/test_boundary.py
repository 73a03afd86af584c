"""CPU: the C-ABI library loads and exports every symbol include/apse_b200.h declares; the product never
touches the oracle; hot calls fail loudly without a GPU."""
import ctypes
import os
import re
import numpy as np
import pytest
from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "apse_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(apse_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from apse_uav_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/apse_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(syms), set(_lib.SIGNATURES) ^ set(syms)
    assert _lib.load().apse_abi_version() == 2


def test_params_struct_matches_c_defaults():
    from apse_uav_b200 import _lib, aruco
    p = _lib.Params()
    _lib.load().apse_params_default(ctypes.byref(p))
    d = aruco.DetectorParameters()
    for name, _ in _lib.Params._fields_:
        assert float(getattr(p, name)) == pytest.approx(float(getattr(d, name)), rel=1e-6), name


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "apse_uav_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower().replace("oracle/cv2", ""), f"{f} mentions the oracle"
                assert not re.search(r"^\s*(import|from)\s+cv2", txt, re.M) or "__getattr__" in txt, f"{f} imports cv2 eagerly"


def test_hot_calls_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import apse_uav_b200 as A
    with pytest.raises(A.ApseError):
        A.cvtColor(np.zeros((8, 8, 3), np.uint8), A.COLOR_BGR2GRAY)
    with pytest.raises(A.ApseError):
        A.aruco.detectMarkers(np.zeros((64, 64), np.uint8), A.aruco.getPredefinedDictionary(A.aruco.DICT_4X4_50))
    with pytest.raises(A.ApseError):
        A.Pipeline(np.eye(3), np.zeros(5), (64, 64), np.arange(256, dtype=np.uint8),
                   A.aruco.getPredefinedDictionary(A.aruco.DICT_4X4_50), None)


def test_embedded_dictionaries_match_cv2():
    cv2 = pytest.importorskip("cv2")
    from apse_uav_b200 import aruco
    for name in ("DICT_4X4_50", "DICT_4X4_250", "DICT_5X5_100", "DICT_6X6_250"):
        a = aruco.getPredefinedDictionary(getattr(aruco, name))
        b = cv2.aruco.getPredefinedDictionary(getattr(cv2.aruco, name))
        assert getattr(aruco, name) == getattr(cv2.aruco, name)
        assert np.array_equal(a.bytesList, b.bytesList) and a.markerSize == b.markerSize
        assert a.maxCorrectionBits == b.maxCorrectionBits
