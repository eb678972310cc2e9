"""C oracle vs live cv2 (plain mode) on random inputs, incl. the 61 MP geometry's
non-integer vertical scale.  Skipped where cv2 is not importable."""
import numpy as np
import pytest

import oracle

cv2 = pytest.importorskip("cv2")


@pytest.fixture(autouse=True)
def plain_mode():
    cv2.setUseOptimized(False)
    yield
    cv2.setUseOptimized(True)


def same(a, b):
    return np.array_equal(np.ascontiguousarray(a, np.float32).view(np.uint32),
                          np.ascontiguousarray(b, np.float32).view(np.uint32))


@pytest.mark.parametrize("w,h", [(211, 123), (640, 480), (1197, 799), (333, 1064)])
def test_sky_chain_and_stats(w, h):
    rng = np.random.default_rng(w * 7 + h)
    g = rng.random((h, w), dtype=np.float32)
    lo = cv2.resize(g, (w // 8, h // 8))
    assert same(oracle.resize_linear(g, w // 8, h // 8), lo)
    md = cv2.medianBlur(lo, 5)
    assert same(oracle.median5(lo), md)
    up = cv2.resize(md, (w, h))
    assert same(oracle.resize_linear(md, w, h), up)
    sky = cv2.GaussianBlur(up, (31, 31), 0)
    assert same(oracle.gauss31(up), sky)
    st = g - sky
    roi = st[h // 20:h // 20 + int(h * 0.9), w // 20:w // 20 + int(w * 0.9)]
    mean, sd = cv2.meanStdDev(roi)
    m2, s2 = oracle.roi_stats(st)
    assert m2 == mean[0, 0] and s2 == sd[0, 0]


@pytest.mark.parametrize("ch", [1, 3])
def test_warp_affine(ch):
    rng = np.random.default_rng(5 + ch)
    h, w = 97, 131
    src = rng.random((h, w, ch), dtype=np.float32) if ch == 3 else rng.random((h, w), dtype=np.float32)
    for _ in range(6):
        ang = rng.uniform(-0.2, 0.2)
        s = rng.uniform(0.8, 1.2)
        M = np.array([[s * np.cos(ang), -s * np.sin(ang), rng.uniform(-30, 30)],
                      [s * np.sin(ang), s * np.cos(ang), rng.uniform(-30, 30)]], np.float32)
        want = cv2.warpAffine(src, M, (w, h), flags=cv2.INTER_LINEAR + cv2.WARP_INVERSE_MAP)
        assert same(oracle.warp_affine(src, M), want)


def test_gray_and_masters():
    rng = np.random.default_rng(9)
    img = rng.random((50, 70, 3), dtype=np.float32)
    assert same(oracle.gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    frames = [rng.random((20, 30, 3), dtype=np.float32) for _ in range(5)]
    bias = rng.random((20, 30, 3), dtype=np.float32) * 0.1
    acc = cv2.subtract(frames[0], bias)
    for f in frames[1:]:
        acc = cv2.add(acc, cv2.subtract(f, bias), dtype=cv2.CV_32F)
    want = acc * np.float32(1.0 / 5)  # Mat /= n is convertTo(.., 1.0/n)
    assert same(oracle.stack_mean(frames, bias), want)
    avg = np.float32(cv2.mean(want)[0])
    got = oracle.flat_normalize(want)
    assert same(got, want * np.float32(1.0 / np.float64(avg)))


def test_calibration_divide_follows_opencv_3_3_on_zero_divisors():
    """getCalibratedImage (util.cpp:269-271): subtract, subtract, divide.  Equal to cv2 wherever the flat is non-zero;
    where it is exactly 0 (either sign) the quotient is 0 — cv::divide of OpenCV 3.3.0, the version the reference pins
    (Dockerfile:7; arithm.cpp div_f: `denom != 0 ? src1 * scale / denom : 0`).  cv2 >= 4 returns inf/nan there, which
    is the one place the oracle deliberately departs from the cv2 4.13 it is otherwise pinned to."""
    rng = np.random.default_rng(10)
    img = rng.random((40, 60, 3), dtype=np.float32)
    bias = rng.random((40, 60, 3), dtype=np.float32) * 0.01
    dark = rng.random((40, 60, 3), dtype=np.float32) * 0.02
    flat = (0.5 + rng.random((40, 60, 3), dtype=np.float32)).astype(np.float32)
    flat[3, 5, 1] = 0.0
    flat[7, 9, 2] = -0.0
    got = oracle.calibrate(img, bias, dark, flat)
    want = cv2.divide(cv2.subtract(cv2.subtract(img, bias), dark), flat)
    nz = flat != 0
    assert same(got[nz], want[nz])
    assert got[3, 5, 1] == 0.0 and got[7, 9, 2] == 0.0 and np.isfinite(got).all()
    assert not np.isfinite(want[3, 5, 1])   # documents the cv2 >= 4 behaviour the reference never saw
