"""CPU: the oracle's pose stage against cv2 and the golden vectors (tolerance 1e-4 relative, BASELINE.json)."""
import numpy as np
import pytest
from conftest import needs_cv2, golden_cases, load_golden

POSE_RTOL = 1e-4


@needs_cv2
def test_project_undistort_rodrigues_vs_cv2(oracle, camera):
    import cv2
    K, D = camera
    rng = np.random.default_rng(0)
    obj = rng.uniform(-3, 3, (56, 3)); obj[:, 2] = 0
    for _ in range(20):
        r = rng.normal(0, 1, 3); t = np.array([rng.uniform(-8, 8), rng.uniform(-5, 5), rng.uniform(20, 50)])
        a, jac = cv2.projectPoints(obj, r, t, K, D)
        b, dr, dt = oracle.project_points(obj, r, t, K, D, jacobian=True)
        assert np.abs(a.reshape(-1, 2) - b).max() < 1e-9
        assert np.abs(jac[:, :3] - dr).max() < 1e-6 * np.abs(jac[:, :3]).max()
        assert np.abs(jac[:, 3:6] - dt).max() < 1e-6 * np.abs(jac[:, 3:6]).max()
        R, J = cv2.Rodrigues(r)
        R2, J2 = oracle.rodrigues(r)
        assert np.abs(R - R2).max() < 1e-14 and np.abs(J - J2).max() < 1e-13
        assert np.abs(cv2.Rodrigues(R)[0].ravel() - oracle.rodrigues_inv(R)).max() < 1e-12
    pts = rng.uniform([200, 200], [3600, 1900], (100, 2))
    a = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D).reshape(-1, 2)
    assert np.abs(a - oracle.undistort_points(pts, K, D)).max() < 1e-13


@needs_cv2
def test_solvepnp_vs_cv2(oracle, camera):
    import cv2
    from oracle import cv2_compat as C
    K, D = camera
    rng = np.random.default_rng(0)
    L = 0.55
    h = np.float32(L) / np.float32(2)
    objm = np.array([[-h, h, 0], [h, h, 0], [h, -h, 0], [-h, -h, 0]], np.float64)
    worst = 0
    for i in range(120):
        yaw = rng.uniform(-np.pi, np.pi); tilt = rng.normal(0, 0.05, 2)
        R = cv2.Rodrigues(np.array([tilt[0], tilt[1], 0.]))[0] @ cv2.Rodrigues(np.array([0, 0, yaw]))[0]
        r0 = cv2.Rodrigues(R)[0].ravel(); t0 = np.array([rng.uniform(-10, 10), rng.uniform(-6, 6), rng.uniform(20, 50)])
        img = (cv2.projectPoints(objm, r0, t0, K, D)[0].reshape(4, 2) + rng.normal(0, 0.1, (4, 2))).astype(np.float32)
        rv, tv, _ = C.estimatePoseSingleMarkers([img.reshape(1, 4, 2)], L, K, D)
        orv, otv = oracle.estimate_pose_single_markers(img.reshape(1, 4, 2), L, K, D)
        worst = max(worst, np.linalg.norm(rv - orv) / np.linalg.norm(rv), np.linalg.norm(tv - otv) / np.linalg.norm(tv))
    assert worst < POSE_RTOL


@pytest.mark.parametrize("name", golden_cases())
def test_pose_vs_golden(oracle, name):
    g = load_golden(name)
    if len(g["ids"]) == 0:
        pytest.skip("no markers in this fixture")
    rv, tv = oracle.estimate_pose_single_markers(g["corners"], 0.55, g["K"], g["D"])
    assert np.abs(rv[:, 0] - g["rvec"]).max() < POSE_RTOL * np.abs(g["rvec"]).max()
    assert np.abs(tv[:, 0] - g["tvec"]).max() < POSE_RTOL * np.abs(g["tvec"]).max()
