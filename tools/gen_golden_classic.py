"""Generates tests/golden/classic_*.npz: outputs of the reference's dependency (cv2 4.13 ArucoDetector) in the CLASSIC
candidate mode (cornerRefinementMethod NONE and SUBPIX; north_star stages 2-4, BASELINE.json config 5) on small
seeded synthetic gray frames, plus per-window adaptiveThreshold checksums and contour counts.  These pin the oracle
(CPU tests) and the CUDA path (GPU tests) without cv2 or /root/reference.  Run here: python tools/gen_golden_classic.py"""
import os, sys, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from oracle import cv2_compat as C
from tools import synth

out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
d = C.Dictionary_get(C.DICT_4X4_50)
cases = [  # name, W, H, seed, synth kwargs, window sweep
    ("classic_dense_960x540", 960, 540, 202, dict(ids=list(range(30)), side_range=(30, 80), jitter=0.15, occlude_frac=0.15, margin=0, noise_sigma=4), (3, 23, 10)),
    ("classic_sparse_640x360", 640, 360, 201, dict(ids=(1, 2, 3, 4), side_range=(36, 60), margin=40), (3, 23, 10)),
    ("classic_odd_643x361", 643, 361, 203, dict(ids=(5, 6, 7, 8, 9), side_range=(40, 70), margin=0, noise_sigma=3), (5, 21, 4)),
    ("classic_empty_320x240", 320, 240, 204, dict(ids=()), (3, 23, 10)),
]
for name, W, H, seed, kw, wins in cases:
    gray = cv2.cvtColor(synth.make_frame(d.bytesList, seed, W, H, **kw), cv2.COLOR_BGR2GRAY)
    rec = dict(gray=gray, wins=np.array(wins), cv2_version=cv2.__version__)
    crcs, ncont = [], []
    for win in range(wins[0], wins[1] + 1, wins[2]):
        w2 = win + 1 if win % 2 == 0 else win
        b = cv2.adaptiveThreshold(gray, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY_INV, w2, 7)
        crcs.append(zlib.crc32(b.tobytes()))
        ncont.append(len(cv2.findContours(b, cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)[0]))
    rec["thresh_crc"] = np.array(crcs, np.uint32); rec["n_contours"] = np.array(ncont)
    for refine, tag in ((C.CORNER_REFINE_NONE, "none"), (C.CORNER_REFINE_SUBPIX, "subpix")):
        p = C.reference_parameters(refine)
        p.adaptiveThreshWinSizeMin, p.adaptiveThreshWinSizeMax, p.adaptiveThreshWinSizeStep = wins
        c, i, r = cv2.aruco.ArucoDetector(d, p).detectMarkers(gray)
        rec[f"ids_{tag}"] = i.ravel() if i is not None else np.zeros(0, np.int32)
        rec[f"corners_{tag}"] = np.array(c, np.float32).reshape(-1, 4, 2)
        rec[f"rejected_{tag}"] = np.array(r, np.float32).reshape(-1, 4, 2)
    np.savez_compressed(os.path.join(out, name + ".npz"), **rec)
    print(name, "markers", len(rec["ids_none"]), "rejected", len(rec["rejected_none"]), "contours", ncont,
          os.path.getsize(os.path.join(out, name + ".npz")) // 1024, "KiB")
