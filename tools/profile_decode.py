"""Development aid: cycles per phase of k_decode on a dense 4K frame of the classic path (library built by `make decprof`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("APSE_LIB", os.path.join(ROOT, "apse_uav_b200", "libapse_b200_decprof.so"))
sys.path.insert(0, ROOT)
import numpy as np, torch
import apse_uav_b200 as A
from apse_uav_b200 import aruco
import __graft_entry__ as G
import bench
from tools import synth

K, D = bench.load_camera()
d = aruco.getPredefinedDictionary(aruco.DICT_4X4_50)
frames = torch.from_numpy(np.stack([synth.make_dense_frame(d.bytesList, 11)])).cuda()
p = G.reference_parameters(aruco); p.cornerRefinementMethod = 1
pipe = A.Pipeline(K, D, (3840, 2160), G.gamma_lut(), d, p, max_batch=1, max_markers=512)
det = pipe.run_batch(frames); torch.cuda.synchronize()
print("second run", flush=True)
det = pipe.run_batch(frames); torch.cuda.synchronize()
print("markers", det["n"].tolist())
