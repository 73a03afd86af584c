"""GPU: seeded 4K frames the other tests do not use (different ids, marker sizes, two dense frames) through the 3-stream
overlapped Pipeline -- every frame against the CPU oracle chain, and three runs against each other (the chain uses
atomics for its work lists and hash tables; results must not depend on their order)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_overlapped_pipeline_matches_oracle_and_is_deterministic(oracle, camera, lut, dictionary, ref_params):
    pytest.importorskip("cv2")   # tools.synth renders the frames with cv2
    import torch
    import apse_uav_b200 as A
    from tools import synth
    K, D = camera
    W, H = 3840, 2160
    frames = []
    for i in range(10):
        if i % 5 == 4:
            frames.append(synth.make_dense_frame(dictionary.bytesList, 500 + i))
        else:
            frames.append(synth.make_frame(dictionary.bytesList, 7000 + 13 * i, W, H, ids=(1, 2, 3, 4, 11, 23)[: 4 + i % 3], side_range=(40, 110)))
    frames = np.stack(frames)
    pipe = A.Pipeline(K, D, (W, H), lut, dictionary, ref_params, max_batch=len(frames), max_markers=256, streams=3, ring=4)
    dev = torch.from_numpy(frames).cuda()
    runs = [A.Pipeline.to_host(pipe.run_batch(dev, want_rejected=True, sync=False, input_ready=True)) for _ in range(3)]
    pipe.close()
    for k in ("n", "ids", "corners", "n_rejected", "rejected", "rvec", "tvec"):
        for r in runs[1:]:
            assert np.array_equal(r[k], runs[0][k]), k
    res = runs[0]
    mx, my = oracle.init_undistort_map(K, D, W, H)
    for i in range(len(frames)):
        _, gray = oracle.preprocess(frames[i], mx, my, lut)
        oc, oi, orj = oracle.detect_markers_apriltag(gray, dictionary.raw, ref_params)
        n = int(res["n"][i])
        assert n == len(oi) and n >= 4 and np.array_equal(res["ids"][i, :n], oi)          # ids and order bit-exact
        assert np.abs(res["corners"][i, :n] - oc).max() <= 1e-3                            # sub-pixel corners, 1e-3 px
        assert int(res["n_rejected"][i]) == len(orj)                                        # candidate counts
        orv, otv = oracle.estimate_pose_single_markers(oc, 0.55, K, D)
        rel = np.linalg.norm(res["tvec"][i, :n] - otv[:, 0], axis=-1) / np.linalg.norm(otv[:, 0], axis=-1)
        assert rel.max() < 1e-4                                                             # pose, 1e-4 relative
