"""GPU parity: stage 1 (aruco_detect.py:250-259,568,592) through the C ABI vs the oracle / cv2 / golden vectors.
Integer pipeline => bit-exact."""
import zlib
import numpy as np
import pytest
from conftest import golden_cases, load_golden, has_cv2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from apse_uav_b200.engine import Engine
    e = Engine(0, 3840, 2160, 4)
    yield e
    e.close()


def test_undistort_map_4k(eng, oracle, camera):
    K, D = camera
    mx, my = eng.init_undistort_map(K, D, 3840, 2160)
    ox, oy = oracle.init_undistort_map(K, D, 3840, 2160)
    assert np.array_equal(mx.cpu().numpy(), ox) and np.array_equal(my.cpu().numpy(), oy)
    if has_cv2():
        import cv2
        cx, cy = cv2.initUndistortRectifyMap(K, D, None, K, (3840, 2160), 5)
        assert ((mx.cpu().numpy() != cx) | (my.cpu().numpy() != cy)).sum() <= 4


@pytest.mark.parametrize("kind", ["sparse", "dense"])
def test_fused_preprocess_4k_bit_exact(eng, oracle, camera, lut, frames4k, kind):
    import torch
    K, D = camera
    eng.set_camera(K, D, 3840, 2160)
    eng.set_lut(lut)
    ox, oy = oracle.init_undistort_map(K, D, 3840, 2160)
    ref_bgr, ref_gray = oracle.preprocess(frames4k[kind], ox, oy, lut)
    out, gray = eng.preprocess(torch.from_numpy(frames4k[kind]).cuda(), want_bgr=True)
    assert np.array_equal(out.cpu().numpy(), ref_bgr)
    assert np.array_equal(gray.cpu().numpy(), ref_gray)
    # batch == per-frame, gray-only path identical
    both = torch.from_numpy(np.stack([frames4k["sparse"], frames4k["dense"], frames4k[kind]])).cuda()
    _, g3 = eng.preprocess(both)
    assert np.array_equal(g3[2].cpu().numpy(), ref_gray)
    if has_cv2():
        import cv2
        from oracle import cv2_compat as C
        mx, my = cv2.initUndistortRectifyMap(K, D, None, K, (3840, 2160), 5)
        ref = C.preprocess_frame(frames4k[kind], mx, my, lut.reshape(1, 256))
        assert np.array_equal(out.cpu().numpy(), ref)
        assert np.array_equal(gray.cpu().numpy(), cv2.cvtColor(ref, cv2.COLOR_BGR2GRAY))


def test_colour_kernels_all_colours(eng, oracle):
    import torch
    c = np.arange(1 << 24, dtype=np.uint32)
    cols = np.stack([(c >> 16) & 255, (c >> 8) & 255, c & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
    t = torch.from_numpy(cols).cuda()
    assert np.array_equal(eng.cvt(t, "rgb2lab").cpu().numpy(), oracle.rgb2lab(cols))
    assert np.array_equal(eng.cvt(t, "lab2rgb").cpu().numpy(), oracle.lab2rgb(cols))
    assert np.array_equal(eng.cvt(t, "bgr2gray").cpu().numpy(), oracle.bgr2gray(cols))


def test_remap_edges_and_ragged_sizes(eng, oracle):
    import torch
    rng = np.random.default_rng(2)
    for (w, h, cn) in [(641, 359, 3), (37, 19, 1), (1280, 720, 1), (8, 8, 3)]:
        src = rng.integers(0, 256, (h, w, cn) if cn == 3 else (h, w), dtype=np.uint8)
        # maps that run off every border (BORDER_CONSTANT 0) with random sub-pixel offsets
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        mx = (xx * 1.13 - 9.37 + rng.uniform(-1, 1, (h, w))).astype(np.float32)
        my = (yy * 1.21 - 7.77 + rng.uniform(-1, 1, (h, w))).astype(np.float32)
        got = eng.remap(torch.from_numpy(src).cuda(), mx, my).cpu().numpy()
        assert np.array_equal(got, oracle.remap(src, mx, my)), (w, h, cn)


@pytest.mark.parametrize("name", golden_cases())
def test_preprocess_golden(name):
    import torch
    from apse_uav_b200.engine import Engine
    g = load_golden(name)
    h, w = g["frame"].shape[:2]
    e = Engine(0, w, h, 1)
    e.set_camera(g["K"], g["D"], w, h)
    e.set_lut(g["lut"])
    out, gray = e.preprocess(torch.from_numpy(g["frame"]).cuda(), want_bgr=True)
    assert zlib.crc32(out.cpu().numpy().tobytes()) == int(g["corrected_crc"])
    assert zlib.crc32(gray.cpu().numpy().tobytes()) == int(g["gray_crc"])
    e.close()


def test_dropin_functions_match_reference_call_sequence(camera, lut, frames4k, oracle):
    """aruco_detect.py:568,250-259,592 with only the import swapped."""
    import apse_uav_b200 as cv2
    K, D = camera
    frame = frames4k["sparse"]
    mapx, mapy = cv2.initUndistortRectifyMap(K, D, None, K, (3840, 2160), 5)
    lookUpTable = lut.reshape(1, 256)
    f = cv2.remap(frame, mapx, mapy, cv2.INTER_LINEAR)
    lab = cv2.cvtColor(f, cv2.COLOR_RGB2LAB)
    lab[..., 0] = cv2.LUT(lab[..., 0], lookUpTable)
    f = cv2.cvtColor(lab, cv2.COLOR_LAB2RGB)
    gray = cv2.cvtColor(f, cv2.COLOR_BGR2GRAY)
    ox, oy = oracle.init_undistort_map(K, D, 3840, 2160)
    ref_bgr, ref_gray = oracle.preprocess(frame, ox, oy, lut)
    assert isinstance(gray, np.ndarray) and np.array_equal(f, ref_bgr) and np.array_equal(gray, ref_gray)
    with pytest.raises(cv2.error):
        cv2.cvtColor(f, 40)  # BGR2HSV is not on the hot path
    with pytest.raises(cv2.error):
        cv2.remap(frame, mapx, mapy, cv2.INTER_NEAREST)


def test_undistort_dropin_bit_exact(oracle, camera, frames4k):
    """cv2.undistort drop-in (apse_undistort): FP64 coordinate -> Q5, bit-exact with the oracle (pinned against cv2.undistort
    on the CPU) and with cv2 itself when importable; 3-channel 4K, 1-channel with a new camera matrix, ragged size."""
    import apse_uav_b200 as A
    K, D = camera
    frame = frames4k["sparse"]
    out = A.undistort(frame, K, D)
    assert isinstance(out, np.ndarray) and np.array_equal(out, oracle.undistort(frame, K, D))
    Ks = K.copy(); Ks[:2] *= 0.25
    newK = Ks.copy(); newK[0, 0] *= 0.9; newK[1, 1] *= 0.95; newK[0, 2] += 7.3
    rng = np.random.default_rng(4)
    g = rng.integers(0, 256, (541, 963), dtype=np.uint8)
    assert np.array_equal(A.undistort(g, Ks, D, None, newK), oracle.undistort(g, Ks, D, newK))
    if has_cv2():
        import cv2
        assert np.array_equal(out, cv2.undistort(frame, K, D))
        assert np.array_equal(A.undistort(g, Ks, D[:5]), cv2.undistort(g, Ks, D[:5]))


def test_fused_kernel_all_colours(oracle, lut):
    """The colour chain INSIDE the TMA-staged fused kernel (its own composed tables and arithmetic, not the stand-alone
    cvtColor kernels) on every 8-bit colour: identity camera (fx = fy = 1, cx = cy = 0, no distortion -> the undistort map is
    the identity and the bilinear remap copies the pixel), three 3840x2160 frames hold all 2^24 colours.  Both output variants
    (gray only = the hot path with the compile-time width, gray + corrected BGR) against the oracle chain."""
    import torch
    from apse_uav_b200.engine import Engine
    W, H = 3840, 2160
    c = np.arange(3 * W * H, dtype=np.uint32) & 0xFFFFFF
    cols = np.stack([(c >> 16) & 255, (c >> 8) & 255, c & 255], -1).astype(np.uint8).reshape(3, H, W, 3)
    assert len(np.unique(c)) == 1 << 24
    e = Engine(0, W, H, 3)
    e.set_camera(np.eye(3), np.zeros(14), W, H)
    e.set_lut(lut)
    t = torch.from_numpy(cols).cuda()
    out, gray = e.preprocess(t, want_bgr=True)
    _, gray_only = e.preprocess(t)
    flat = cols.reshape(-1, W, 3)
    lab = oracle.rgb2lab(flat)
    lab[..., 0] = lut[lab[..., 0]]
    ref = oracle.lab2rgb(lab)
    assert np.array_equal(out.cpu().numpy().reshape(ref.shape), ref)
    ref_gray = oracle.bgr2gray(ref)
    assert np.array_equal(gray.cpu().numpy().reshape(ref_gray.shape), ref_gray)
    assert np.array_equal(gray_only.cpu().numpy().reshape(ref_gray.shape), ref_gray)
    e.close()


@pytest.mark.parametrize("size", [(1080, 720), (1000, 564), (644, 360)])
def test_fused_preprocess_widths_without_tma(oracle, camera, lut, size):
    """Frame widths whose row pitch (3 w bytes) is not a multiple of 16 cannot be described by a TMA tensor map: the
    batch takes the generic fused kernel instead of failing, with the same bit-exact result."""
    import torch
    from apse_uav_b200.engine import Engine
    K, D = camera
    w, h = size
    Ks = K.copy(); Ks[:2] *= w / 3840.0
    rng = np.random.default_rng(w)
    frames = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
    e = Engine(0, w, h, 2)
    e.set_camera(Ks, D, w, h)
    e.set_lut(lut)
    out, gray = e.preprocess(torch.from_numpy(frames).cuda(), want_bgr=True)
    ox, oy = oracle.init_undistort_map(Ks, D, w, h)
    for i in range(2):
        rb, rg = oracle.preprocess(frames[i], ox, oy, lut)
        assert np.array_equal(out[i].cpu().numpy(), rb) and np.array_equal(gray[i].cpu().numpy(), rg)
    e.close()
