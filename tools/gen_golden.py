"""Generates tests/golden/*.npz from the reference's own implementation (cv2 4.13 through oracle/cv2_compat.py,
i.e. the call sequence of aruco_detect.py:250-269,592-601) on small seeded synthetic frames.  These vectors pin
the oracle (CPU tests) and the CUDA path (GPU tests); /root/reference is not needed to consume them.
Run here (cv2 required):  python tools/gen_golden.py"""
import os, sys, json, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from oracle import cv2_compat as C
from tools import synth

out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
cam = json.load(open(os.path.join(out, "cam_params.json")))
K0 = np.array(cam["mtx"]); D = np.array(cam["dist"]).ravel()
d = C.Dictionary_get(C.DICT_4X4_50)
lut = C.gamma_lut()
cases = [  # name, W, H, seed, kwargs
    ("sparse_640x360", 640, 360, 101, dict(ids=(1, 2, 3, 4), side_range=(36, 60), margin=50)),
    ("dense_960x540", 960, 540, 102, dict(ids=list(range(24)), side_range=(30, 70), jitter=0.15, occlude_frac=0.15, margin=0, noise_sigma=4)),
    ("odd_643x361", 643, 361, 103, dict(ids=(5, 6, 7), side_range=(40, 60), margin=50)),
    ("empty_320x240", 320, 240, 104, dict(ids=())),
]
for name, W, H, seed, kw in cases:
    K = K0.copy(); K[:2] *= W / 3840.0
    frame = synth.make_frame(d.bytesList, seed, W, H, **kw)
    mapx, mapy = cv2.initUndistortRectifyMap(K, D, None, K, (W, H), 5)
    r = C.reference_chain(frame, mapx, mapy, lut, C.reference_parameters(), K, D)
    n = 0 if r["ids"] is None else len(r["ids"])
    np.savez_compressed(os.path.join(out, name + ".npz"), frame=frame, K=K, D=D, lut=lut.ravel(),
                        gray_crc=np.uint32(zlib.crc32(r["gray"].tobytes())), corrected_crc=np.uint32(zlib.crc32(r["corrected"].tobytes())),
                        gray_rows=r["gray"][::37].copy(),
                        ids=(r["ids"].ravel() if n else np.zeros(0, np.int32)), corners=np.array(r["corners"], np.float32).reshape(-1, 4, 2),
                        rejected=np.array(r["rejected"], np.float32).reshape(-1, 4, 2),
                        rvec=(r["rvec"].reshape(-1, 3) if n else np.zeros((0, 3))), tvec=(r["tvec"].reshape(-1, 3) if n else np.zeros((0, 3))),
                        cv2_version=cv2.__version__)
    print(name, "markers", n, "rejected", len(r["rejected"]), os.path.getsize(os.path.join(out, name + ".npz")) // 1024, "KiB")
