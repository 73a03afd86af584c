"""GPU parity: estimatePoseSingleMarkers / projectPoints (aruco_detect.py:601,344,377,424,468); 1e-4 relative."""
import numpy as np
import pytest
from conftest import golden_cases, load_golden, has_cv2

pytestmark = pytest.mark.gpu
POSE_RTOL = 1e-4


def test_pose_vs_oracle_and_cv2(oracle, camera):
    from apse_uav_b200 import aruco
    K, D = camera
    rng = np.random.default_rng(0)
    L = 0.55
    h = np.float32(L) / np.float32(2)
    objm = np.array([[-h, h, 0], [h, h, 0], [h, -h, 0], [-h, -h, 0]], np.float64)
    corners = []
    for i in range(300):
        yaw = rng.uniform(-np.pi, np.pi); tilt = rng.normal(0, 0.05, 2)
        rv = np.array([tilt[0], tilt[1], 0.]); R1, _ = oracle.rodrigues(rv); R2, _ = oracle.rodrigues(np.array([0, 0, yaw]))
        r0 = oracle.rodrigues_inv(R1 @ R2); t0 = np.array([rng.uniform(-10, 10), rng.uniform(-6, 6), rng.uniform(20, 50)])
        img = oracle.project_points(objm, r0, t0, K, D) + rng.normal(0, 0.1, (4, 2))
        corners.append(img.astype(np.float32).reshape(1, 4, 2))
    rv, tv, obj = aruco.estimatePoseSingleMarkers(corners, L, K, D)
    assert rv.shape == (300, 1, 3) and rv.dtype == np.float64 and obj.shape == (4, 1, 3) and obj.dtype == np.float32
    orv, otv = oracle.estimate_pose_single_markers(np.array(corners), L, K, D)
    rr = np.linalg.norm(rv - orv, axis=-1) / np.linalg.norm(orv, axis=-1)
    rt = np.linalg.norm(tv - otv, axis=-1) / np.linalg.norm(otv, axis=-1)
    assert rr.max() < POSE_RTOL and rt.max() < POSE_RTOL
    assert np.median(rr) < 1e-9 and np.median(rt) < 1e-9
    if has_cv2():
        from oracle import cv2_compat as C
        crv, ctv, _ = C.estimatePoseSingleMarkers(corners, L, K, D)
        assert (np.linalg.norm(rv - crv, axis=-1) / np.linalg.norm(crv, axis=-1)).max() < POSE_RTOL
        assert (np.linalg.norm(tv - ctv, axis=-1) / np.linalg.norm(ctv, axis=-1)).max() < POSE_RTOL


def test_project_points(oracle, camera):
    import apse_uav_b200 as cv2
    K, D = camera
    rng = np.random.default_rng(1)
    obj = rng.uniform(-3, 3, (56, 3)); obj[:, 2] = 0
    for _ in range(5):
        r = rng.normal(0, 1, 3); t = np.array([rng.uniform(-8, 8), rng.uniform(-5, 5), rng.uniform(20, 50)])
        img, jac = cv2.projectPoints(obj, r, t, K, D)
        assert img.shape == (56, 1, 2) and jac is None
        assert np.abs(img.reshape(-1, 2) - oracle.project_points(obj, r, t, K, D)).max() < 1e-8
    img, _ = cv2.projectPoints(np.float32([[0, 0.42, 0]]), r.reshape(1, 3), t.reshape(1, 3), K, D)   # aruco_detect.py:377 shapes
    assert img.shape == (1, 1, 2)


@pytest.mark.parametrize("name", golden_cases())
def test_pose_golden(name):
    from apse_uav_b200 import aruco
    g = load_golden(name)
    if len(g["ids"]) == 0:
        rv, tv, _ = aruco.estimatePoseSingleMarkers([], 0.55, g["K"], g["D"])
        assert rv.shape == (0, 1, 3)
        return
    rv, tv, _ = aruco.estimatePoseSingleMarkers(tuple(c.reshape(1, 4, 2) for c in g["corners"]), 0.55, g["K"], g["D"])
    assert np.abs(rv[:, 0] - g["rvec"]).max() < POSE_RTOL * np.abs(g["rvec"]).max()
    assert np.abs(tv[:, 0] - g["tvec"]).max() < POSE_RTOL * np.abs(g["tvec"]).max()
