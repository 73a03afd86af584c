"""CPU: the host post-pass (aruco_detect.py:598-782) against the reference script's own CSV rows
(tests/golden/sequence_4k.json, produced by running aruco_detect.py through tools/run_reference_script.py)."""
import json
import os
import numpy as np
import pytest
from conftest import GOLDEN, needs_cv2


def _golden_rows(name="sequence_4k.json"):
    g = json.load(open(os.path.join(GOLDEN, name)))
    rows = [[float(v) for v in line.split(",")] for line in g["csv"][1:]]
    return g, np.array(rows)


@pytest.fixture(scope="module")
def sequence_records(oracle, camera, lut, dictionary, ref_params):
    """per-frame detection records of the golden sequence from the ORACLE chain, pose at the nominal marker length."""
    pytest.importorskip("cv2")  # the frame renderer (tools.synth) needs cv2 for warpPerspective / resize
    from tools import synth
    g, _ = _golden_rows()
    K, D = camera
    ox, oy = oracle.init_undistort_map(K, D, 3840, 2160)
    recs = []
    for k, f in enumerate(synth.make_sequence(dictionary.bytesList, g["base_seed"], g["n_frames"])):
        _, gray = oracle.preprocess(f, ox, oy, lut)
        c, ids, _ = oracle.detect_markers_apriltag(gray, dictionary.raw, ref_params)
        rv, tv = oracle.estimate_pose_single_markers(c, 0.55, K, D)
        recs.append(dict(frame=k, ids=ids, corners=c, rvec=rv[:, 0], tvec=tv[:, 0]))
    return recs


def test_yaw_matches_scipy():
    from scipy.spatial.transform import Rotation as R
    from apse_uav_b200.postpass import yaw_zxy_deg
    rng = np.random.default_rng(0)
    for _ in range(200):
        r = rng.normal(0, 1.5, 3)
        assert abs(yaw_zxy_deg(r) - R.from_rotvec(r).as_euler("zxy", degrees=True)[0]) < 1e-9


def test_two_pass_sequence_matches_reference_csv(sequence_records, oracle, camera):
    from apse_uav_b200 import shard
    K, D = camera
    _, ref = _golden_rows()
    project = lambda obj, r, t: oracle.project_points(obj, r, t, K, D)
    lengths = shard.scan_marker_lengths(sequence_records, project)
    recs = shard.exact_pose(sequence_records, lengths,
                            lambda c, ml: tuple(np.concatenate(x) for x in zip(*[
                                (lambda rv, tv: (rv[:, 0], tv[:, 0]))(*oracle.estimate_pose_single_markers(c[i:i + 1], float(ml[i]), K, D))
                                for i in range(len(c))])))
    rows = shard.final_scan(recs, project)
    from apse_uav_b200.postpass import csv_line
    from conftest import assert_csv_rows_match
    # values are rounded to 2-5 decimals in the CSV: one unit in the last printed place of each column on top of 1e-4 relative
    assert_csv_rows_match(rows, ref)
    assert len(csv_line(rows[0]).split(",")) == 16


def test_gating_marks_new_marker_as_false_positive_for_one_frame(oracle, camera):
    """:609-613,636-637: a marker that was absent on the previous frame is 'detected' but flagged (id -> -1) and takes no
    part in distances until the next frame; missing frames leave detected_ID_prev untouched (:782 is inside the if)."""
    from apse_uav_b200.postpass import SequencePostPass
    K, D = camera
    project = lambda obj, r, t: oracle.project_points(obj, r, t, K, D)
    pp = SequencePostPass(project)
    sq = lambda cx, cy, s=30: np.float32([[cx - s, cy - s], [cx + s, cy - s], [cx + s, cy + s], [cx - s, cy + s]])
    def pose(c):
        rv, tv = oracle.estimate_pose_single_markers(c, pp.marker_length, K, D)
        return rv[:, 0], tv[:, 0]
    c = np.stack([sq(1000, 1000), sq(1600, 1100)])
    r1 = pp.step(1, np.array([4, 1]), c, *pose(c))
    assert r1["ID_4_detected"] == 1 and r1["ID_1_detected"] == 1 and r1["distance_veh1_aruco"] > 0
    c2 = np.stack([sq(1001, 1000), sq(1601, 1100), sq(2200, 900)])
    r2 = pp.step(2, np.array([4, 1, 2]), c2, *pose(c2))
    assert r2["ID_2_detected"] == 1 and r2["distance_veh2_aruco"] == 0       # new marker: flagged, no distance yet
    r3 = pp.step(3, None, np.zeros((0, 4, 2)), np.zeros((0, 3)), np.zeros((0, 3)))
    assert r3["ID_4_detected"] == 0 and pp.detected_prev == [1, 1, 0, 1]
    c4 = np.stack([sq(1002, 1000), sq(1900, 1100), sq(2201, 900)])          # marker 1 jumped 300 px: gated out
    r4 = pp.step(4, np.array([4, 1, 2]), c4, *pose(c4))
    assert r4["ID_1_detected"] == 0 and r4["ID_2_detected"] == 1 and r4["distance_veh2_aruco"] > 0


def test_led_readout_matches_reference_csv(oracle, camera, lut, dictionary, ref_params):
    """aruco_detect.py:338-373: the LED strip of the host vehicle, read back from the corrected gray frame
    with the numpy expression of the reference; golden = the reference script's rows for a sequence with rendered LEDs."""
    pytest.importorskip("cv2")
    from tools import synth
    from apse_uav_b200 import shard
    from apse_uav_b200.postpass import CSV_FIELDS
    g, ref = _golden_rows("sequence_4k_leds.json")
    assert sorted(set(ref[:, 3].astype(int))) == [0, 77, 102, 129, 178, 255]      # the patterns that were drawn
    K, D = camera
    ox, oy = oracle.init_undistort_map(K, D, 3840, 2160)
    recs, grays = [], []
    for k, f in enumerate(synth.make_sequence(dictionary.bytesList, g["base_seed"], g["n_frames"], leds=g["leds"])):
        _, gray = oracle.preprocess(f, ox, oy, lut)
        c, ids, _ = oracle.detect_markers_apriltag(gray, dictionary.raw, ref_params)
        rv, tv = oracle.estimate_pose_single_markers(c, 0.55, K, D)
        recs.append(dict(frame=k, ids=ids, corners=c, rvec=rv[:, 0], tvec=tv[:, 0]))
        grays.append(gray)
    project = lambda obj, r, t: oracle.project_points(obj, r, t, K, D)
    lengths = shard.scan_marker_lengths(recs, project)
    recs = shard.exact_pose(recs, lengths, lambda c, ml: tuple(np.concatenate(x) for x in zip(*[
        (lambda rv, tv: (rv[:, 0], tv[:, 0]))(*oracle.estimate_pose_single_markers(c[i:i + 1], float(ml[i]), K, D)) for i in range(len(c))])))
    led_mean = lambda frame: (lambda x, y: np.sum(np.sum(grays[frame][y - 2:y + 3, x - 2:x + 3])) / 25)
    rows = shard.final_scan(recs, project, led_mean_for_frame=led_mean)
    from conftest import assert_csv_rows_match
    assert_csv_rows_match(rows, ref)                                                             # incl. leds_ID


def test_events_sequence_matches_reference_csv(oracle, camera, lut, dictionary, ref_params):
    """64 frames in which a vehicle vanishes and returns, a vehicle and the host jump further than DIFF_MAX, the host
    vanishes (altitude fallback, aruco_detect.py:639-642) and one frame has no marker at all (:599): the gating / relabel
    branches (:613,637,669) and every stale value of the CSV, through BOTH host implementations of the loop -- the Python
    mirror (postpass.py) and the native scan (csrc/sequence.cu) -- against the reference script's own rows."""
    pytest.importorskip("cv2")
    from tools import synth
    from apse_uav_b200 import shard, sequence
    from conftest import assert_csv_rows_match, golden_csv, golden_events
    from test_sequence_native import eval_jobs_numpy
    g, ref = golden_csv("sequence_4k_events.json")
    flags = ref[:, [7, 10, 13, 1]]
    assert (flags == 0).any() and (ref[40] == [41] + [0] * 15).all()
    K, D = camera
    ox, oy = oracle.init_undistort_map(K, D, 3840, 2160)
    recs, grays = [], []
    for k, f in enumerate(synth.make_sequence(dictionary.bytesList, g["base_seed"], g["n_frames"], leds=g["leds"], events=golden_events(g))):
        _, gray = oracle.preprocess(f, ox, oy, lut)
        c, ids, _ = oracle.detect_markers_apriltag(gray, dictionary.raw, ref_params)
        rv, tv = oracle.estimate_pose_single_markers(c, 0.55, K, D) if len(ids) else (np.zeros((0, 1, 3)), np.zeros((0, 1, 3)))
        recs.append(dict(frame=k, ids=ids, corners=c, rvec=rv[:, 0], tvec=tv[:, 0]))
        grays.append(gray)
    project = lambda obj, r, t: oracle.project_points(obj, r, t, K, D)
    pose_each = lambda c, ml: tuple(np.concatenate(x) for x in zip(*[
        (lambda rv, tv: (rv[:, 0], tv[:, 0]))(*oracle.estimate_pose_single_markers(c[i:i + 1], float(ml[i]), K, D)) for i in range(len(c))]))
    lengths = shard.scan_marker_lengths(recs, project)
    recs2 = shard.exact_pose(recs, lengths, pose_each)
    led_mean = lambda frame: (lambda x, y: np.sum(np.sum(grays[frame][y - 2:y + 3, x - 2:x + 3])) / 25)
    rows = shard.final_scan(recs2, project, led_mean_for_frame=led_mean)
    assert_csv_rows_match(rows, ref)
    # the native scan on the same records: identical rows
    F, M = len(recs), 8
    n = np.array([len(r["ids"]) for r in recs2], np.int32)
    ids = np.full((F, M), -1, np.int32); corners = np.zeros((F, M, 4, 2), np.float32); rvec = np.zeros((F, M, 3)); tvec = np.zeros((F, M, 3))
    for k, r in enumerate(recs2):
        ids[k, :n[k]], corners[k, :n[k]], rvec[k, :n[k]], tvec[k, :n[k]] = r["ids"], r["corners"].reshape(-1, 4, 2), r["rvec"], r["tvec"]
    _, nrows, jobs = sequence.scan(sequence.seq_config(leds=True), n, ids, corners, rvec, tvec)
    res = eval_jobs_numpy(jobs, project, lambda k, x, y: led_mean(k)(x, y))
    assert sequence.rows_to_dicts(sequence.finish(nrows, res)) == rows
