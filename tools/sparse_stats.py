"""Development aid: fraction of 4x4 tiles the sparse evaluation computes exactly, per frame, on the bench frames."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
import apse_uav_b200 as A
from apse_uav_b200 import aruco
import __graft_entry__ as G
K, D = bench.load_camera()
d = aruco.getPredefinedDictionary(aruco.DICT_4X4_50)
B = int(os.environ.get("B", "8"))
pipe = A.Pipeline(K, D, (3840, 2160), G.gamma_lut(), d, G.reference_parameters(aruco), max_batch=B, max_markers=64)
e = pipe.engine
one = bench.base_frames(1)
for name, fr in (("same frame x B", np.repeat(one, B, 0)), ("bench_sequence", bench.base_sequence(min(B, 4)))):
    frames = torch.from_numpy(fr).cuda()
    n_f = frames.shape[0]
    gray = torch.empty(frames.shape[:3], dtype=torch.uint8, device="cuda")
    e.preprocess_tiles(frames, gray, sparse=True)
    flags = torch.zeros((n_f, 540, 960), dtype=torch.uint8, device="cuda")
    n = C.c_int(0)
    e._check(e.lib.apse_debug_sparse(e.h, None, flags.data_ptr(), n_f, C.byref(n), e._stream()))
    print(name, "exact tile fraction per frame", flags.float().mean((1, 2)).cpu().numpy().round(4), "total", n.value)
