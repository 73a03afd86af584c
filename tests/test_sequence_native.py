"""CPU: the native sequence post-pass (csrc/sequence.cu: apse_sequence_scan / _finish / _csv, host code) against the
per-frame Python mirror of aruco_detect.py:598-782 (apse_uav_b200/postpass.py) on fabricated detection records with
drop-outs, jumps beyond DIFF_MAX, re-appearances, duplicated and foreign ids.  The device half (apse_sequence_jobs) is
replaced here by a numpy evaluation of the same jobs through the oracle's projectPoints; tests/test_gpu_pipeline.py runs the
real kernel."""
import numpy as np
import pytest

from apse_uav_b200 import postpass, sequence, shard
from apse_uav_b200._lib import SEQ_RESULT_DTYPE


def fabricate(oracle, K, D, F=160, M=8, seed=0):
    """Detection records of a synthetic flight: ids 1-4 drift, vanish, jump and come back; id 7 and a duplicate appear."""
    rng = np.random.default_rng(seed)
    pos = {4: np.array([1000., 1000.]), 1: np.array([1700., 1150.]), 2: np.array([2300., 800.]), 3: np.array([2600., 1500.])}
    vel = {i: rng.normal(0, 1.2, 2) for i in pos}
    sq = lambda c, s, a: (np.array([[-s, -s], [s, -s], [s, s], [-s, s]]) @ np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]).T + c)
    n = np.zeros(F, np.int32)
    ids = np.full((F, M), -1, np.int32)
    corners = np.zeros((F, M, 4, 2), np.float32)
    for k in range(F):
        present = []
        for i in (4, 1, 2, 3):
            pos[i] = pos[i] + vel[i] + rng.normal(0, 0.2, 2)
            if rng.random() < 0.03:
                pos[i] = pos[i] + rng.normal(0, 150, 2)            # a jump far beyond DIFF_MAX
            gone = rng.random() < (0.10 if i != 4 else 0.06)
            if not gone:
                present.append(i)
        order = list(rng.permutation(present))
        if rng.random() < 0.08:
            order.append(7)                                          # foreign id
        if rng.random() < 0.04 and present:
            order.append(present[0])                                 # the same id twice in one frame
        for j, i in enumerate(order[:M]):
            c = pos[i] if i in pos else np.array([500., 1800.])
            if j >= len(present) and i in pos:
                c = c + np.array([300., -200.])
            corners[k, j] = sq(c, 30 + (i % 3) + rng.normal(0, 0.3), rng.uniform(-0.5, 0.5)).astype(np.float32)
            ids[k, j] = i
        n[k] = min(len(order), M)
    rvec = np.zeros((F, M, 3))
    tvec = np.zeros((F, M, 3))
    for k in range(F):
        if n[k]:
            rv, tv = oracle.estimate_pose_single_markers(corners[k, :n[k]], 0.55, K, D)
            rvec[k, :n[k]], tvec[k, :n[k]] = rv[:, 0], tv[:, 0]
    return n, ids, corners, rvec, tvec


def eval_jobs_numpy(jobs, project, gray_mean=None):
    """apse_sequence_jobs restated with numpy (projection through `project`)."""
    res = np.zeros(len(jobs), SEQ_RESULT_DTYPE)
    for i, J in enumerate(jobs):
        if J["kind"] == 0:
            px = postpass.to_pixels(project(postpass.LED_AXIS, J["rvec"], J["tvec"]))
            leds = 0
            for j in range(8):
                if gray_mean(int(J["frame"]), int(px[j][0]), int(px[j][1])) > J["led_threshold"]:
                    leds += 2 ** (7 - j)
            res[i]["leds"] = leds
        else:
            px = postpass.to_pixels(project(postpass.bbox_points(J["dim"]), J["rvec"], J["tvec"]))
            d = np.sqrt((np.float64(J["src"][0]) - px[:, 0]) ** 2 + (np.float64(J["src"][1]) - px[:, 1]) ** 2)
            t = int(np.argmin(d))
            a = np.sqrt((J["src"][0] - J["tgt"][0]) * (J["src"][0] - J["tgt"][0]) + (J["src"][1] - J["tgt"][1]) * (J["src"][1] - J["tgt"][1]))
            b = np.sqrt((J["src"][0] - px[t][0]) * (J["src"][0] - px[t][0]) + (J["src"][1] - px[t][1]) * (J["src"][1] - px[t][1]))
            res[i]["dist_aruco"], res[i]["dist_bbox"] = float(a * J["scale"]), float(b * J["scale"])
        res[i]["valid"] = 1
    return res


def python_rows(n, ids, corners, rvec, tvec, project, led_mean=None):
    pp = postpass.SequencePostPass(project, start_frame=1)
    rows, lengths = [], []
    for k in range(len(n)):
        lengths.append(pp.marker_length)
        if led_mean is not None:
            pp.led_mean = lambda x, y, k=k: led_mean(k, x, y)
        m = int(n[k])
        rows.append(pp.step(1 + k, ids[k, :m] if m else None, corners[k, :m], rvec[k, :m], tvec[k, :m]))
    return rows, lengths


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_native_scan_equals_python_mirror(oracle, camera, seed):
    K, D = camera
    n, ids, corners, rvec, tvec = fabricate(oracle, K, D, seed=seed)
    project = lambda obj, r, t: oracle.project_points(obj, r, t, K, D)
    # deterministic stand-in for the gray frames of the LED read-out
    led_mean = lambda k, x, y: float((k * 37 + x * 11 + y * 7) % 256)
    want, want_len = python_rows(n, ids, corners, rvec, tvec, project, led_mean)
    cfg = sequence.seq_config(start_frame=1, leds=True)
    lengths, rows, jobs = sequence.scan(cfg, n, ids, corners, rvec, tvec)
    assert np.array_equal(lengths, np.array(want_len)), "marker length recurrence differs"
    res = eval_jobs_numpy(jobs, project, led_mean)
    rows = sequence.finish(rows, res)
    got = sequence.rows_to_dicts(rows)
    assert got == want
    # the branches the golden sequences never reach (aruco_detect.py:613,637,669): gated-out markers and re-appearances
    det = np.array([[r[f"ID_{v}_detected"] for v in (1, 2, 3)] + [r["ID_4_detected"]] for r in want])
    assert (det == 0).any() and (det[1:] != det[:-1]).any()
    # CSV text: Python's str() of every field
    assert sequence.rows_to_csv(rows) == shard.rows_to_csv(want)


def test_yaw_sign_shortcut_at_the_rounding_boundary(oracle, camera):
    """aruco_detect.py:412-414 uses the vehicle's yaw only as `round(yaw, 2) < 0`.  The native scan decides that from the two
    matrix entries the angle is the arctangent of and evaluates the angle itself only near |yaw| = 0.005 degrees; here: rotation
    vectors about the optical axis whose yaw is the rotation angle, on and around the boundary, zero, minus zero and half turns --
    the outline dimensions the distance jobs carry must equal the Python mirror's."""
    K, D = camera
    project = lambda obj, r, t: oracle.project_points(obj, r, t, K, D)
    deg = [0.0, -0.0, 0.004, -0.004, -0.005, -0.0050001, -0.0049999, -0.006, -0.015, -0.5, -0.58, 0.58, 90.0, -90.0, 179.9999, -179.9999,
           180.0, -180.0, 1e-9, -1e-9, 33.3, -33.3]
    F, M = len(deg), 4
    n = np.full(F, 2, np.int32)
    ids = np.full((F, M), -1, np.int32)
    ids[:, 0], ids[:, 1] = 4, 1
    sq = np.array([[-15., -15.], [15., -15.], [15., 15.], [-15., 15.]])
    corners = np.zeros((F, M, 4, 2), np.float32)
    corners[:, 0] = sq + np.array([1000., 1000.])
    corners[:, 1] = sq + np.array([1700., 1150.])
    rvec = np.zeros((F, M, 3))
    tvec = np.zeros((F, M, 3))
    tvec[:, :2] = np.array([[-3.0, -0.5, 30.0], [2.0, 0.7, 30.0]])
    rvec[:, 0, 2] = 0.3
    rvec[:, 1, 2] = np.deg2rad(np.array(deg))
    want, _ = python_rows(n, ids, corners, rvec, tvec, project)
    lengths, rows, jobs = sequence.scan(sequence.seq_config(start_frame=1), n, ids, corners, rvec, tvec)
    assert len(jobs) == F and (jobs["kind"] == 1).all()
    rows = sequence.finish(rows, eval_jobs_numpy(jobs, project))
    assert sequence.rows_to_dicts(rows) == want
    # and the flip itself, from the job's dimensions: dims[1] *= 1 + ah / 2 with ah = atan(2 / 30) > 0, negated unless the yaw
    # rounds to a negative number
    flipped = jobs["dim"][:, 1] / postpass.VEH_DIM[1][1] < 1
    expect_negative = np.array([round(float(np.rad2deg(np.arctan2(np.sin(np.deg2rad(d)), np.cos(np.deg2rad(d))))), 2) < 0 for d in deg])
    assert np.array_equal(flipped, ~expect_negative), (flipped, expect_negative)


def test_native_first_pass_lengths_equal_python(oracle, camera):
    K, D = camera
    n, ids, corners, rvec, tvec = fabricate(oracle, K, D, F=60, seed=5)
    project = lambda obj, r, t: oracle.project_points(obj, r, t, K, D)
    recs = [dict(frame=k, ids=ids[k, :n[k]], corners=corners[k, :n[k]], rvec=rvec[k, :n[k]], tvec=tvec[k, :n[k]]) for k in range(len(n))]
    want = shard.scan_marker_lengths(recs, project)
    lengths, _, _ = sequence.scan(sequence.seq_config(), n, ids, corners, rvec, tvec, rescale_tvec=True, want_rows=False)
    assert np.array_equal(lengths, np.array(want))


def test_csv_float_formatting_matches_python_str():
    rows = np.zeros(6, sequence.SEQ_ROW_DTYPE)
    vals = [0.53467, 26.74, 100.0, 1e-05, 123456.789, 0.001]
    for i, v in enumerate(vals):
        rows[i]["frame_id"] = i + 1
        rows[i]["detected"] = [1, 0, 1, 1]
        rows[i]["host_fields"] = 1
        rows[i]["marker_length"], rows[i]["altitude"], rows[i]["fov_width"], rows[i]["fov_height"] = v, -v, v * 3, 0.0
        rows[i]["dist_aruco"] = [v, 0, 2 * v]
        rows[i]["dist_bbox"] = [0.0, 0, 7.0]
    text = sequence.rows_to_csv(rows, header=False).splitlines()
    for i, v in enumerate(vals):
        want = ",".join(str(x) for x in [i + 1, 1, v, 0, -v, v * 3, 0.0, 1, v, 0.0, 0, 0, 0, 1, 2 * v, 7.0])
        assert text[i] == want


def test_py_round_equals_python_round():
    """The CSV columns of aruco_detect.py:146-185 are Python round(float, n): the native 128-bit integer path equals the
    interpreter on random values, exact decimal ties in binary (k / 2^j) and values one ulp either side of a tie."""
    from apse_uav_b200 import _lib
    f = _lib.load().apse_py_round
    rng = np.random.default_rng(3)
    vals = list(rng.uniform(-4000, 4000, 40000)) + list(rng.uniform(-2, 2, 20000)) + list(rng.normal(0, 1e-3, 5000))
    vals += [k / 8 for k in range(-400, 400)] + [k / 16 + 0.03125 for k in range(-200, 200)] + [k / 64 for k in range(-999, 999)]
    vals += [0.5, 1.5, 2.5, 0.125, 0.375, 2.675, 1.005, 1e-9, -1e-9, 123456.789, 0.0005, 0.00049999999999999, 1e15, 3.9e15, 1e300]
    ties = [k / 1000 + 0.0005 for k in range(0, 3000, 7)]
    vals += ties + [float(np.nextafter(v, np.inf)) for v in ties] + [float(np.nextafter(v, -np.inf)) for v in ties]
    for nd in (0, 2, 3, 5):
        for v in vals:
            assert f(float(v), nd) == round(float(v), nd), (v, nd)


def test_csv_float_fast_path_equals_python_str():
    """apse_sequence_csv prints round(value, n) columns from the integer q of q / 10^n (sequence.cu, py_float_str): equal to
    Python's str() on rounded values of every column width, on values that are NOT short decimals (general path), on
    negative values, tiny values (scientific notation) and integers."""
    rng = np.random.default_rng(11)
    vals = []
    for nd in (0, 1, 2, 3, 5, 6):
        vals += [round(float(v), nd) for v in rng.uniform(-5000, 5000, 3000)]
        vals += [round(float(v), nd) for v in rng.uniform(-1, 1, 1000)]
    vals += [float(v) for v in rng.uniform(-100, 100, 2000)]                      # not short decimals
    vals += [1e-05, 5e-05, 9.9e-05, 0.0001, 0.00011, 1e8, 999999999.0, 1e9, 1234567890.12, 1e15, 1e16, 0.1 + 0.2, 1 / 3, 2.675, 1e-9]
    rows = np.zeros(len(vals), sequence.SEQ_ROW_DTYPE)
    rows["frame_id"] = np.arange(len(vals)) + 1
    rows["detected"][:, 3] = 1
    rows["host_fields"] = 1
    rows["marker_length"] = vals
    text = sequence.rows_to_csv(rows, header=False).splitlines()
    assert len(text) == len(vals)
    for v, line in zip(vals, text):
        assert line.split(",")[2] == str(v), (v, line)


@pytest.mark.parametrize("chunks", [[60], [7, 1, 30, 22], [1] * 60, [59, 1]])
def test_chunked_scan_equals_whole_sequence(oracle, camera, chunks):
    """apse_sequence_scan_chunk / _finish_chunk (a sequence that arrives in pieces: shard.run_sequence_streamed): lengths, rows
    and jobs of both passes, and the finished rows with their stale values, equal the single-call scan for any chunking."""
    K, D = camera
    n, ids, corners, rvec, tvec = fabricate(oracle, K, D, F=60, seed=9)
    project = lambda obj, r, t: oracle.project_points(obj, r, t, K, D)
    cfg = sequence.seq_config()
    len1, _, _ = sequence.scan(cfg, n, ids, corners, rvec, tvec, rescale_tvec=True, want_rows=False)
    len2, rows, jobs = sequence.scan(cfg, n, ids, corners, rvec, tvec)
    want = sequence.finish(rows, eval_jobs_numpy(jobs, project))
    s1, s2 = sequence.new_state(), sequence.new_state()
    got_rows, got_len1, got_len2, lo = [], [], [], 0
    for c in chunks:
        sl = slice(lo, lo + c)
        l1, _, _ = sequence.scan(cfg, n[sl], ids[sl], corners[sl], rvec[sl], tvec[sl], rescale_tvec=True, want_rows=False, state=s1, frame0=lo)
        l2, r, j = sequence.scan(cfg, n[sl], ids[sl], corners[sl], rvec[sl], tvec[sl], state=s2, frame0=lo)
        assert (j["frame"] >= lo).all() and (j["frame"] < lo + c).all()
        got_rows.append(sequence.finish_chunk(s2, r, eval_jobs_numpy(j, project)))
        got_len1.append(l1); got_len2.append(l2)
        lo += c
    assert np.array_equal(np.concatenate(got_len1), len1) and np.array_equal(np.concatenate(got_len2), len2)
    got = np.concatenate(got_rows)
    for f in ("frame_id", "detected", "host_fields", "leds", "accepted_mask", "marker_length", "altitude", "fov_width", "fov_height",
              "dist_aruco", "dist_bbox"):
        assert np.array_equal(got[f], want[f]), f
    assert sequence.rows_to_csv(got) == sequence.rows_to_csv(want)
