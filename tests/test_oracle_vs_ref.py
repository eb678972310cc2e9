"""C oracle vs the reference's own object code (oracle/_ref) on fresh random cases.
Skipped where oracle/_ref was never built (no /root/reference)."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.skipif(oracle.ref() is None, reason="oracle/_ref not built")





def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_hfti_bitwise_2000_systems():
    rng = np.random.default_rng(101)
    for t in range(2000):
        m = int(rng.integers(6, 121))
        a = np.zeros((3, 120), np.float32)
        b = np.zeros((2, 120), np.float32)
        x = rng.integers(0, 9576, m).astype(np.float32)
        y = rng.integers(0, 6388, m).astype(np.float32)
        if t % 3 == 0:
            x, y = y * 0.3, x * 1.0  # make the y column win the pivot
        a[0, :m], a[1, :m], a[2, :m] = x, y, 1
        b[0, :m] = np.round(x + rng.normal(0, 2, m))
        b[1, :m] = np.round(y + rng.normal(0, 2, m))
        ao, bo, ko, ipo = oracle.hfti(a, b, m)
        ar, br, kr, ipr = oracle.hfti(a, b, m, which="ref")
        assert np.array_equal(bits(ao), bits(ar)) and np.array_equal(bits(bo), bits(br))
        assert ko == kr and ipo == ipr


def test_quicksort_matches_and_only_slot0_is_ever_misplaced():
    rng = np.random.default_rng(102)
    stuck = 0
    for t in range(300):
        n = int(rng.integers(0, 400)) if t % 2 else int(rng.integers(2000, 9880))
        x = rng.uniform(0.5, 0.9, n).astype(np.float32)
        if t % 5 == 0 and n > 10:
            x[rng.integers(0, n, n // 3)] = x[0]
        xo, io = oracle.sort_keys(x)
        xr, ir = oracle.sort_keys(x, which="ref")
        assert np.array_equal(xo, xr) and np.array_equal(io, ir)
        if n > 2:
            assert np.all(np.diff(xo[1:]) >= 0)
            stuck += int(xo[0] > xo[1])
    assert stuck > 30


def test_alignment_random_lists():
    rng = np.random.default_rng(103)
    for t in range(120):
        n1 = int(rng.integers(3, 60))
        xy1 = np.stack([rng.integers(0, 6000, n1), rng.integers(0, 800, n1)], 1).astype(np.int32)
        th = rng.uniform(-0.003, 0.003)
        c, s = np.cos(th), np.sin(th)
        sh = rng.uniform(-8, 8, 2)
        xy2 = np.stack([c * xy1[:, 0] - s * xy1[:, 1] + sh[0], s * xy1[:, 0] + c * xy1[:, 1] + sh[1]], 1)
        xy2 = np.floor(xy2 + rng.normal(0, 0.4, (n1, 2))).astype(np.int32)
        i1, t1 = oracle.triangles(xy1)
        i2, t2 = oracle.triangles(xy1, which="ref")
        assert np.array_equal(i1, i2) and np.array_equal(bits(t1), bits(t2))
        o, r = oracle.align(xy1, xy2), oracle.align(xy1, xy2, which="ref")
        assert o["err"] == r["err"] and o["k"] == r["k"]
        assert np.array_equal(o["matches"][:, :o["k"]], r["matches"][:, :r["k"]])
        assert np.array_equal(bits(o["M"]), bits(r["M"]))


def test_detect_deblend_on_blended_blobs():
    total = 0
    rng = np.random.default_rng(104)
    yy, xx = np.mgrid[0:150, 0:200]
    for seed in range(4):
        img = rng.normal(0, 0.002, (150, 200))
        for _ in range(40):
            x, y = rng.uniform(0, 200), rng.uniform(0, 150)
            a = np.exp(rng.uniform(np.log(0.01), np.log(0.9)))
            s = rng.uniform(1.2, 3.5)
            for k in range(int(rng.integers(1, 4))):
                img += a * rng.uniform(0.3, 1) * np.exp(
                    -0.5 * ((xx + 0.5 - x - k * rng.uniform(-6, 6)) ** 2 + (yy + 0.5 - y - k * rng.uniform(-6, 6)) ** 2) / s ** 2)
        img = img.astype(np.float32)
        for th in (0.004, 0.01, 0.05):
            so, nco = oracle.detect(img, th, th * 4 / 3)
            sr, ncr = oracle.detect(img, th, th * 4 / 3, which="ref")
            assert nco == ncr and len(so) == len(sr)
            for f in ("x", "y", "area"):
                assert np.array_equal(so[f], sr[f])
            assert np.array_equal(bits(so["value"]), bits(sr["value"]))
            total += len(so)
    assert total > 500
