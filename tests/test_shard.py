"""CPU, world_size 2 over gloo: frame sharding + gather + sequential scan reproduce the single-process result
(the N > 1 host path of SURVEY.md 8e; the GPU work per rank is replaced by precomputed detection records)."""
import os
import pickle
import subprocess
import sys
import numpy as np
from conftest import ROOT


def test_shard_bounds_cover_everything():
    from apse_uav_b200.shard import shard_bounds
    for n in (0, 1, 7, 1800):
        for w in (1, 2, 3, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def test_round_plan_deals_the_sequence_out_in_order():
    """shard.round_plan (streamed runs): rounds and parts are consecutive, cover the sequence once, stay within the batch size,
    are balanced over the ranks, and the optional last round is short."""
    from apse_uav_b200.shard import round_plan
    for n in (0, 1, 7, 130, 1800, 1801):
        for w in (1, 2, 3, 8):
            for batch, tail in ((60, 0), (60, 8), (6, 2)):
                plan = round_plan(n, w, batch, tail)
                flat = [p for rnd in plan for p in rnd]
                assert all(len(rnd) == w for rnd in plan)
                assert sum(hi - lo for lo, hi in flat) == n
                if flat:
                    assert flat[0][0] == 0 and flat[-1][1] == n and all(flat[i][1] == flat[i + 1][0] for i in range(len(flat) - 1))
                    assert max(hi - lo for lo, hi in flat) <= batch
                for rnd in plan:
                    sizes = [hi - lo for lo, hi in rnd]
                    assert max(sizes) - min(sizes) <= 1
                if tail and n >= 2 * w * batch:
                    assert sum(hi - lo for lo, hi in plan[-1]) == tail * w


WORKER = r'''
import os, pickle, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch.distributed as dist
from apse_uav_b200 import shard
from oracle import oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[3], rank=rank, world_size=world)
recs, K, D = pickle.load(open(sys.argv[2], "rb"))
lo, hi = shard.shard_bounds(len(recs), world)[rank]
allrec = shard.gather_records(recs[lo:hi], rank, world)
if rank == 0:
    project = lambda obj, r, t: O.project_points(obj, r, t, K, D)
    lengths = shard.scan_marker_lengths(allrec, project)
    rows = shard.final_scan(allrec, project)
    pickle.dump((lengths, rows), open(sys.argv[2] + ".out", "wb"))
dist.barrier()
dist.destroy_process_group()
'''


def test_world2_gloo_matches_single_process(tmp_path, oracle, camera):
    from apse_uav_b200 import shard
    K, D = camera
    rng = np.random.default_rng(3)
    sq = lambda cx, cy, s: np.float32([[cx - s, cy - s], [cx + s, cy - s], [cx + s, cy + s], [cx - s, cy + s]])
    recs = []
    pos = {4: [1000., 1000.], 1: [1700., 1150.], 2: [2300., 800.], 3: [2600., 1500.]}
    for k in range(9):
        ids = [4, 1, 2, 3] if k != 4 else [4, 2]          # one frame loses two vehicles
        c = np.stack([sq(pos[i][0] + 0.7 * k, pos[i][1] - 0.4 * k, 30 + (i % 3)) for i in ids]) + rng.normal(0, 0.05, (len(ids), 4, 2)).astype(np.float32)
        rv, tv = oracle.estimate_pose_single_markers(c.astype(np.float32), 0.55, K, D)
        recs.append(dict(frame=k, ids=np.array(ids), corners=c.astype(np.float32), rvec=rv[:, 0], tvec=tv[:, 0]))
    project = lambda obj, r, t: oracle.project_points(obj, r, t, K, D)
    ref_lengths = shard.scan_marker_lengths(recs, project)
    ref_rows = shard.final_scan(recs, project)
    data = tmp_path / "recs.pkl"
    pickle.dump((recs, K, D), open(data, "wb"))
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29600 + os.getpid() % 300)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(data), port],
                              env=dict(os.environ, RANK=str(r), WORLD_SIZE="2")) for r in range(2)]
    assert all(p.wait(timeout=300) == 0 for p in procs)
    lengths, rows = pickle.load(open(str(data) + ".out", "rb"))
    assert lengths == ref_lengths and rows == ref_rows
    assert rows[4]["ID_1_detected"] == 0 and rows[5]["ID_1_detected"] == 1


WORKER_NATIVE = r'''
import os, pickle, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import numpy as np, torch
import torch.distributed as dist
from apse_uav_b200 import shard, sequence
from oracle import oracle as O
from test_sequence_native import eval_jobs_numpy
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[3], rank=rank, world_size=world)
arrs, K, D = pickle.load(open(sys.argv[2], "rb"))
F = len(arrs["n"])
lo, hi = shard.shard_bounds(F, world)[rank]
det = {k: torch.from_numpy(v[lo:hi].copy()) for k, v in arrs.items()}
det["status"] = torch.zeros(hi - lo, dtype=torch.int32)
allv, sizes = shard.gather_detections(det, hi - lo, rank, world)
if rank == 0:
    assert sizes == [b - a for a, b in shard.shard_bounds(F, world)]
    h = {k: v.numpy() for k, v in allv.items()}
    project = lambda obj, r, t: O.project_points(obj, r, t, K, D)
    lengths, rows, jobs = sequence.scan(sequence.seq_config(), h["n"], h["ids"], h["corners"], h["rvec"], h["tvec"])
    rows = sequence.finish(rows, eval_jobs_numpy(jobs, project))
    pickle.dump((lengths, sequence.rows_to_dicts(rows)), open(sys.argv[2] + ".out", "wb"))
dist.barrier()
dist.destroy_process_group()
'''


def test_world3_gloo_tensor_gather_and_native_scan(tmp_path, oracle, camera):
    """the N > 1 host path as shard.run_sequence runs it: unequal blocks, results gathered as padded tensors (dist.gather),
    native scan on rank 0 == the same scan on the unsharded arrays"""
    from apse_uav_b200 import sequence
    from test_sequence_native import fabricate, eval_jobs_numpy
    K, D = camera
    n, ids, corners, rvec, tvec = fabricate(oracle, K, D, F=50, seed=11)
    project = lambda obj, r, t: oracle.project_points(obj, r, t, K, D)
    ref_len, rows, jobs = sequence.scan(sequence.seq_config(), n, ids, corners, rvec, tvec)
    ref_rows = sequence.rows_to_dicts(sequence.finish(rows, eval_jobs_numpy(jobs, project)))
    data = tmp_path / "arrs.pkl"
    pickle.dump((dict(n=n, ids=ids, corners=corners, rvec=rvec, tvec=tvec), K, D), open(data, "wb"))
    script = tmp_path / "worker_native.py"
    script.write_text(WORKER_NATIVE)
    port = str(29900 + os.getpid() % 90)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(data), port],
                              env=dict(os.environ, RANK=str(r), WORLD_SIZE="3")) for r in range(3)]
    assert all(p.wait(timeout=300) == 0 for p in procs)
    lengths, got = pickle.load(open(str(data) + ".out", "rb"))
    assert np.array_equal(lengths, ref_len) and got == ref_rows
