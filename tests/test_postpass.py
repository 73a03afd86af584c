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
    from apse_uav_b200.postpass import CSV_FIELDS, csv_line
    got = np.array([[float(r[f]) for f in CSV_FIELDS] for r in rows])
    assert got.shape == ref.shape
    assert np.array_equal(got[:, [0, 1, 3, 7, 10, 13]], ref[:, [0, 1, 3, 7, 10, 13]])        # frame ids, detection flags, leds
    num = [2, 4, 5, 6, 8, 9, 11, 12, 14, 15]
    # values are rounded to 2-5 decimals in the CSV: allow one unit in the last printed place on top of 1e-4 relative
    assert np.all(np.abs(got[:, num] - ref[:, num]) <= 1e-4 * np.abs(ref[:, num]) + 0.0101)
    assert np.abs(got[:, 2] - ref[:, 2]).max() <= 1.01e-5                                        # markerLength (5 decimals)
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
    got = np.array([[float(r[f]) for f in CSV_FIELDS] for r in rows])
    assert np.array_equal(got[:, [0, 1, 3, 7, 10, 13]], ref[:, [0, 1, 3, 7, 10, 13]])        # incl. leds_ID
    num = [2, 4, 5, 6, 8, 9, 11, 12, 14, 15]
    assert np.all(np.abs(got[:, num] - ref[:, num]) <= 1e-4 * np.abs(ref[:, num]) + 0.0101)
