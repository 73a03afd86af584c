"""GPU: the batched device-resident Pipeline (preprocess -> detect -> pose) vs the per-frame oracle chain, plus
size-independent properties at the full 4K size (BASELINE.json configs 2-4)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pipe(camera, lut, dictionary, ref_params):
    import apse_uav_b200 as A
    K, D = camera
    p = A.Pipeline(K, D, (3840, 2160), lut, dictionary, ref_params, max_batch=6, max_markers=256)
    yield p
    p.engine.close()


def test_pipeline_matches_oracle_chain(pipe, oracle, camera, lut, dictionary, ref_params, frames4k):
    import torch
    import apse_uav_b200 as A
    K, D = camera
    ox, oy = oracle.init_undistort_map(K, D, 3840, 2160)
    frames = [frames4k["sparse"], frames4k["dense"]]
    res = A.Pipeline.to_host(pipe.run(torch.from_numpy(np.stack(frames)).cuda()))
    for i, f in enumerate(frames):
        _, gray = oracle.preprocess(f, ox, oy, lut)
        oc, oi, _ = oracle.detect_markers_apriltag(gray, dictionary.raw, ref_params)
        orv, otv = oracle.estimate_pose_single_markers(oc, 0.55, K, D)
        n = int(res["n"][i])
        assert n == len(oi) and np.array_equal(res["ids"][i, :n], oi)
        assert np.abs(res["corners"][i, :n] - oc).max() <= 1e-3
        assert (np.linalg.norm(res["tvec"][i, :n] - otv[:, 0], axis=-1) / np.linalg.norm(otv[:, 0], axis=-1)).max() < 1e-4
        assert (np.linalg.norm(res["rvec"][i, :n] - orv[:, 0], axis=-1) / np.linalg.norm(orv[:, 0], axis=-1)).max() < 1e-4


def test_sequence_properties_full_size(pipe, dictionary, frames4k):
    """14 frames > max_batch (chunking), translated copies of one frame: every frame finds the same ids, corners
    move by the translation (away from the border, where undistortion is locally a shift-invariant warp only
    approximately -> 2 px bound), results independent of the position inside the batch, launches are counted."""
    import torch
    import apse_uav_b200 as A
    base = torch.from_numpy(frames4k["sparse"]).cuda()
    shifts = [(0, 0), (3, 5), (0, 0), (-4, 2), (7, -6), (0, 0), (1, 1), (0, 0), (2, -3), (-5, -5), (0, 0), (4, 4), (6, 0), (0, 0)]
    seq = torch.stack([torch.roll(base, shifts=(dy, dx), dims=(0, 1)) for dx, dy in shifts])
    l0 = pipe.launches
    res = A.Pipeline.to_host(pipe.run(seq))
    assert pipe.launches - l0 >= 3 * 12      # 3 batches x (1 preprocess + 10 detect + 1 decode + 1 pose)
    n0 = int(res["n"][0])
    assert n0 >= 4
    ref_ids = res["ids"][0, :n0]
    zero = [i for i, s in enumerate(shifts) if s == (0, 0)]
    for i in zero:   # identical frames -> bit-identical results wherever they sit in a batch
        assert np.array_equal(res["ids"][i], res["ids"][0]) and np.array_equal(res["corners"][i], res["corners"][0])
        assert np.array_equal(res["tvec"][i, :n0], res["tvec"][0, :n0])
    for i, (dx, dy) in enumerate(shifts):
        n = int(res["n"][i])
        assert sorted(res["ids"][i, :n].tolist()) == sorted(ref_ids.tolist())
        for m in range(n):
            j = list(ref_ids).index(res["ids"][i, m])
            d = res["corners"][i, m] - res["corners"][0, j] - np.float32([dx, dy])
            assert np.abs(d).max() < 2.0
    # pose sanity: tvec_z within the 15-60 m band the synthetic marker sizes imply
    assert ((res["tvec"][0, :n0, 2] > 10) & (res["tvec"][0, :n0, 2] < 80)).all()


def test_sequence_csv_matches_reference_script(pipe, dictionary):
    """End to end on the GPU: frames -> Pipeline -> two-pass post-pass -> CSV rows of aruco_detect.py:146-185, compared
    with the rows the reference script itself wrote for the same seeded frames (tests/golden/sequence_4k.json)."""
    import json, os, torch
    pytest.importorskip("cv2")   # tools.synth renders the frames with cv2.warpPerspective / resize
    from conftest import GOLDEN
    from tools import synth
    from apse_uav_b200 import shard
    from apse_uav_b200.postpass import CSV_FIELDS
    g = json.load(open(os.path.join(GOLDEN, "sequence_4k.json")))
    ref = np.array([[float(v) for v in line.split(",")] for line in g["csv"][1:]])
    frames = torch.from_numpy(np.stack(list(synth.make_sequence(dictionary.bytesList, g["base_seed"], g["n_frames"])))).cuda()
    rows = shard.run_sequence(pipe, frames)
    got = np.array([[float(r[f]) for f in CSV_FIELDS] for r in rows])
    assert got.shape == ref.shape
    assert np.array_equal(got[:, [0, 1, 3, 7, 10, 13]], ref[:, [0, 1, 3, 7, 10, 13]])
    num = [2, 4, 5, 6, 8, 9, 11, 12, 14, 15]
    assert np.all(np.abs(got[:, num] - ref[:, num]) <= 1e-4 * np.abs(ref[:, num]) + 0.0101)
    assert np.abs(got[:, 2] - ref[:, 2]).max() <= 1.01e-5


def test_multi_stream_equals_single_stream(pipe, camera, lut, dictionary, ref_params, frames4k):
    """sub-batches in flight on several streams / contexts give bit-identical results."""
    import torch
    import apse_uav_b200 as A
    K, D = camera
    p3 = A.Pipeline(K, D, (3840, 2160), lut, dictionary, ref_params, max_batch=5, max_markers=256, streams=3, ring=6)
    frames = torch.from_numpy(np.stack([frames4k["sparse"], frames4k["dense"], frames4k["sparse"], frames4k["dense"], frames4k["sparse"]])).cuda()
    a = A.Pipeline.to_host(p3.run_batch(frames, want_rejected=True))
    b = A.Pipeline.to_host(pipe.run_batch(frames, want_rejected=True))
    for k in ("n", "ids", "corners", "n_rejected", "rejected", "rvec", "tvec"):
        assert np.array_equal(a[k], b[k]), k
    # overlapped mode: results are waited for later; consecutive batches share contexts and staging buffers
    frames2 = torch.roll(frames, shifts=(5, 3), dims=(1, 2))
    pending = [p3.run_batch(f, want_rejected=True, sync=False, input_ready=True) for f in (frames, frames2, frames, frames2, frames)]
    outs = [A.Pipeline.to_host(d) for d in pending]
    b2 = A.Pipeline.to_host(pipe.run_batch(frames2, want_rejected=True))
    for o, ref in zip(outs, (b, b2, b, b2, b)):
        for k in ("n", "ids", "corners", "n_rejected", "rejected", "rvec", "tvec"):
            assert np.array_equal(o[k], ref[k]), k
    p3.close()


def test_host_stream_equals_device_batches(pipe, frames4k):
    """end-to-end host path (double-buffered H2D overlapped with compute) returns what run_batch returns"""
    import torch
    f = torch.from_numpy(np.stack([frames4k["sparse"], frames4k["dense"], frames4k["sparse"]]))
    host = [f[:2].clone().pin_memory(), f[1:].clone().pin_memory(), f[:1].clone().pin_memory()]
    outs = list(pipe.run_host_stream(host))
    assert len(outs) == 3
    for hb, o in zip(host, outs):
        ref = pipe.to_host(pipe.run_batch(hb.cuda()))
        for k in ("n", "ids", "corners", "rvec", "tvec"):
            assert np.array_equal(o[k], ref[k]), k


def test_patch_sums_numpy_slicing(pipe):
    """LED read-out kernel = np.sum(gray[y-2:y+3, x-2:x+3]) including numpy's negative-start / clamped-stop slicing"""
    import torch
    rng = np.random.default_rng(9)
    gray = rng.integers(0, 256, (2, 37, 53), dtype=np.uint8)
    pts = [(f, x, y) for f in (0, 1) for x in (0, 1, 2, 3, 25, 49, 50, 51, 52, 53, 60) for y in (0, 1, 2, 17, 33, 34, 35, 36, 37, 40)]
    got = pipe.engine.patch_sums(torch.from_numpy(gray).cuda(), pts, half=2)
    ref = [int(np.sum(gray[f][y - 2:y + 3, x - 2:x + 3])) for f, x, y in pts]
    assert got.tolist() == ref


def test_sequence_with_leds_matches_reference_script(pipe, dictionary):
    """frames with a rendered LED strip -> leds_ID column of the reference script's CSV"""
    import json, os, torch
    pytest.importorskip("cv2")
    from conftest import GOLDEN
    from tools import synth
    from apse_uav_b200 import shard
    from apse_uav_b200.postpass import CSV_FIELDS
    g = json.load(open(os.path.join(GOLDEN, "sequence_4k_leds.json")))
    ref = np.array([[float(v) for v in line.split(",")] for line in g["csv"][1:]])
    frames = torch.from_numpy(np.stack(list(synth.make_sequence(dictionary.bytesList, g["base_seed"], g["n_frames"], leds=g["leds"])))).cuda()
    rows = shard.run_sequence(pipe, frames, leds=True)
    got = np.array([[float(r[f]) for f in CSV_FIELDS] for r in rows])
    assert np.array_equal(got[:, [0, 1, 3, 7, 10, 13]], ref[:, [0, 1, 3, 7, 10, 13]])
    assert got[:, 3].astype(int).tolist() == g["leds"]
    num = [2, 4, 5, 6, 8, 9, 11, 12, 14, 15]
    assert np.all(np.abs(got[:, num] - ref[:, num]) <= 1e-4 * np.abs(ref[:, num]) + 0.0101)
