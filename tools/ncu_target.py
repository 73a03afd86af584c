"""Small fixed workload for ncu captures: 3 pipeline steps over a batch of B sparse synthetic 4K frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import apse_uav_b200 as A
from apse_uav_b200 import aruco
import __graft_entry__ as G
import bench
from tools import synth

B = int(os.environ.get("B", "16")); dense = os.environ.get("DENSE", "0") == "1"
K, D = bench.load_camera()
d = aruco.getPredefinedDictionary(aruco.DICT_4X4_50)
pipe = A.Pipeline(K, D, (3840, 2160), G.gamma_lut(), d, G.reference_parameters(aruco), max_batch=B, max_markers=256 if dense else 64)
base = [synth.make_dense_frame(d.bytesList, 11 + i) if dense else synth.make_frame(d.bytesList, 1000 + i) for i in range(2)]
frames = torch.from_numpy(np.stack([base[i % 2] for i in range(B)])).cuda()
for _ in range(int(os.environ.get("STEPS", "3"))):
    det = pipe.run_batch(frames)
torch.cuda.synchronize()
print("markers", det["n"][:4].tolist(), "launches", pipe.launches)
