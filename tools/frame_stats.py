"""Work statistics of the bench frames (active CCL tiles, boundary points, clusters, quads) -- sizing aid."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import apse_uav_b200 as A
from apse_uav_b200 import aruco
import __graft_entry__ as G

K, D = bench.load_camera()
d = aruco.getPredefinedDictionary(aruco.DICT_4X4_50)
pipe = A.Pipeline(K, D, (bench.W, bench.H), G.gamma_lut(), d, G.reference_parameters(aruco), max_batch=1)
e = pipe.engine
frames = bench.base_frames(3)
for fr in frames:
    _, gray = e.preprocess(torch.from_numpy(fr).cuda())
    r = e.debug_apriltag(gray)
    t = r["thresh"].cpu().numpy()
    act = (t != 127)
    tiles = act.reshape(bench.H // 16, 16, bench.W // 32, 32).any(axis=(1, 3))
    print(dict(active_px=int(act.sum()), active_ccl_tiles=int(tiles.sum()), ccl_tiles=tiles.size, points=r["points"],
               clusters=r["clusters"], fitted=r["fitted"], quads=r["n_quads"]))
