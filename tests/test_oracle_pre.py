"""CPU: the oracle's preprocess stage against the reference's dependency (cv2) and the golden vectors."""
import zlib
import numpy as np
import pytest
from conftest import needs_cv2, golden_cases, load_golden


@needs_cv2
def test_undistort_map_vs_cv2(oracle, camera):
    import cv2
    K, D = camera
    mx, my = cv2.initUndistortRectifyMap(K, D, None, K, (3840, 2160), 5)
    ox, oy = oracle.init_undistort_map(K, D, 3840, 2160)
    bad = (ox != mx) | (oy != my)
    assert bad.sum() <= 4  # cv2 accumulates along the row; <= 1 float32 ulp on a handful of pixels
    assert np.abs(ox - mx).max() <= 5e-4 and np.abs(oy - my).max() <= 5e-4
    # and those pixels map to the same Q5 fixed-point coordinate, i.e. remap output is unaffected
    assert np.array_equal(np.rint(ox[bad] * np.float32(32)), np.rint(mx[bad] * np.float32(32)))


@needs_cv2
def test_remap_bit_exact(oracle, camera):
    import cv2
    K, D = camera
    K = K.copy(); K[:2] *= 0.5
    mx, my = cv2.initUndistortRectifyMap(K, D, None, K, (1920, 1080), 5)
    img = np.random.default_rng(0).integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
    assert np.array_equal(cv2.remap(img, mx, my, cv2.INTER_LINEAR), oracle.remap(img, mx, my))
    g = img[..., 0].copy()
    assert np.array_equal(cv2.remap(g, mx, my, cv2.INTER_LINEAR), oracle.remap(g, mx, my))


@needs_cv2
def test_colour_chain_all_colours(oracle, lut):
    import cv2
    c = np.arange(1 << 24, dtype=np.uint32)
    cols = np.stack([(c >> 16) & 255, (c >> 8) & 255, c & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
    lab = cv2.cvtColor(cols, cv2.COLOR_RGB2LAB)
    assert np.array_equal(lab, oracle.rgb2lab(cols))
    assert np.array_equal(cv2.cvtColor(cols, cv2.COLOR_LAB2RGB), oracle.lab2rgb(cols))
    assert np.array_equal(cv2.cvtColor(cols, cv2.COLOR_BGR2GRAY), oracle.bgr2gray(cols))
    lab[..., 0] = cv2.LUT(lab[..., 0], lut.reshape(1, 256))
    ref = cv2.cvtColor(lab, cv2.COLOR_LAB2RGB)
    olab = oracle.rgb2lab(cols)
    olab[..., 0] = lut[olab[..., 0]]
    assert np.array_equal(ref, oracle.lab2rgb(olab))


@pytest.mark.parametrize("name", golden_cases())
def test_preprocess_vs_golden(oracle, name):
    g = load_golden(name)
    h, w = g["frame"].shape[:2]
    mx, my = oracle.init_undistort_map(g["K"], g["D"], w, h)
    out, gray = oracle.preprocess(g["frame"], mx, my, g["lut"])
    assert zlib.crc32(out.tobytes()) == int(g["corrected_crc"])
    assert zlib.crc32(gray.tobytes()) == int(g["gray_crc"])
    assert np.array_equal(gray[::37], g["gray_rows"])


@needs_cv2
def test_undistort_equals_cv2_undistort(oracle, camera):
    """cv2.undistort rounds the FP64 source coordinate to the Q5 grid directly (CV_16SC2 maps): not the float32-map remap"""
    import cv2
    K, D = camera
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (2160, 3840, 3), dtype=np.uint8)
    ref = cv2.undistort(img, K, D)
    assert np.array_equal(oracle.undistort(img, K, D), ref)
    mx, my = oracle.init_undistort_map(K, D, 3840, 2160)
    assert 0 < (oracle.remap(img, mx, my) != ref).sum() < 0.01 * ref.size     # the two variants really differ
    Ks = K.copy(); Ks[:2] *= 0.25
    newK = Ks.copy(); newK[0, 0] *= 0.9; newK[1, 1] *= 0.95; newK[0, 2] += 7.3
    g = rng.integers(0, 256, (540, 960), dtype=np.uint8)
    assert np.array_equal(oracle.undistort(g, Ks, D, newK), cv2.undistort(g, Ks, D, None, newK))
    assert np.array_equal(oracle.undistort(g, Ks, D[:5]), cv2.undistort(g, Ks, D[:5]))


@needs_cv2
def test_quad_image_resize_and_blur_bit_exact(oracle):
    """aprilTagQuadDecimate / aprilTagQuadSigma (aruco_detect.py:203,231-233): the image the quad detector sees = cv2.resize(
    INTER_AREA) by an integer factor, then cv2.GaussianBlur / unsharp masking with floor(4 |sigma|) | 1 taps."""
    import cv2
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (360, 480), dtype=np.uint8)
    for shape in ((360, 480), (361, 483), (725, 1283)):
        im = rng.integers(0, 256, shape, dtype=np.uint8)
        for f in (1.5, 2, 2.5, 3, 1.3, 4, 5, 6, 8, 1.1, 3.7):     # aruco_detect.py:203 names 1.5
            fx = float(np.float32(1) / np.float32(f))                # the dependency divides in float32
            want = cv2.resize(im, None, fx=fx, fy=fx, interpolation=cv2.INTER_AREA)
            got, scale = oracle.quad_image(im, float(f), 0.0)
            assert got.shape == want.shape and np.array_equal(got, want), (shape, f)
            assert scale == np.float32(f)
    for sigma in list(np.arange(0.3, 4.0, 0.1)) + [0.8, 1.3, 2.3, 5.0, 7.9]:
        s = float(np.float32(sigma))
        ksz = int(np.floor(4 * np.float32(sigma))) | 1
        blur = cv2.GaussianBlur(img, (ksz, ksz), s, sigmaY=s, borderType=cv2.BORDER_REPLICATE) if ksz > 1 else img
        assert np.array_equal(oracle.quad_image(img, 0.0, s)[0], blur), sigma
        sharp = np.clip(2 * img.astype(int) - blur, 0, 255).astype(np.uint8) if ksz > 1 else img
        assert np.array_equal(oracle.quad_image(img, 0.0, -s)[0], sharp), -sigma
