"""Regenerates tests/golden/*.npz.  Run in the authoring container only:

    python tests/golden/make_golden.py

golden_cv2.npz  inputs + outputs of the OpenCV primitives on the hot path, taken
                from cv2 (4.13.0) in PLAIN mode (cv2.setUseOptimized(False)), the
                mode the oracle is pinned to.
golden_ref.npz  inputs + outputs of the reference's own object code (oracle/_ref:
                focas.cpp, adjoiningpixel.cpp and the prebuilt Fortran HFTI driven
                through oracle/ref_harness.cpp).
The reference's test-suite has no numeric vectors for this path,
so these files are the committed pins; tests/test_oracle_golden.py checks the
C oracle against them bit for bit, with neither cv2 nor /root/reference needed.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cv2  # noqa: E402

import oracle  # noqa: E402


def blob_image(w, h, n, seed):
    r = np.random.default_rng(seed)
    img = r.normal(0, 0.002, (h, w))
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(n):
        x, y = r.uniform(0, w), r.uniform(0, h)
        a = np.exp(r.uniform(np.log(0.01), np.log(0.9)))
        s = r.uniform(1.2, 3.5)
        img += a * np.exp(-0.5 * ((xx + 0.5 - x) ** 2 + (yy + 0.5 - y) ** 2) / s ** 2)
        if r.random() < 0.5:
            x2, y2 = x + r.uniform(-8, 8), y + r.uniform(-8, 8)
            img += a * r.uniform(0.3, 1) * np.exp(-0.5 * ((xx + 0.5 - x2) ** 2 + (yy + 0.5 - y2) ** 2) / s ** 2)
    return img.astype(np.float32)


def star_lists(rng, t):
    n1 = int(rng.integers(3, 60))
    xy1 = np.stack([rng.integers(0, 6000, n1), rng.integers(0, 800, n1)], 1).astype(np.int32)
    th = rng.uniform(-0.003, 0.003)
    sh = rng.uniform(-8, 8, 2)
    c, s = np.cos(th), np.sin(th)
    xy2 = np.stack([c * xy1[:, 0] - s * xy1[:, 1] + sh[0], s * xy1[:, 0] + c * xy1[:, 1] + sh[1]], 1)
    xy2 = np.floor(xy2 + rng.normal(0, 0.4, (n1, 2))).astype(np.int32)
    if t % 2:
        keep = rng.random(n1) > 0.15
        extra = np.stack([rng.integers(0, 6000, 5), rng.integers(0, 800, 5)], 1).astype(np.int32)
        xy2 = np.concatenate([xy2[keep], extra])[rng.permutation(int(keep.sum()) + 5)]
    return xy1, xy2


def make_cv2():
    cv2.setUseOptimized(False)
    rng = np.random.default_rng(7)
    out = {}
    u = rng.integers(0, 65536, (24, 40, 3), dtype=np.uint16)
    out["u16"] = u
    f = np.empty(u.shape, np.float32)
    tmp = cv2.multiply(u.astype(np.float32), np.float32(1 / 65535.0).item(), dtype=cv2.CV_32F)
    out["u16_f32"] = tmp
    bgr = rng.random((40, 56, 3), dtype=np.float32)
    out["bgr"] = bgr
    out["bgr_gray"] = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    # sky chain on an odd-sized frame with non-integer vertical scale (like 6388/798)
    w, h = 203, 165
    g = (rng.random((h, w), dtype=np.float32) * 0.01 + 0.05).astype(np.float32)
    g[40:44, 100:105] += 0.5
    out["gray"] = g
    lo = cv2.resize(g, (w // 8, h // 8))
    md = cv2.medianBlur(lo, 5)
    up = cv2.resize(md, (w, h))
    sky = cv2.GaussianBlur(up, (31, 31), 0)
    out.update(sky_lo=lo, sky_med=md, sky_up=up, sky=sky)
    st = g - sky
    roi = st[h // 20:h // 20 + int(h * 0.9), w // 20:w // 20 + int(w * 0.9)]
    mean, sd = cv2.meanStdDev(roi)
    out["roi_mean_std"] = np.array([mean[0, 0], sd[0, 0]], np.float64)
    out["gauss_kernel"] = cv2.getGaussianKernel(31, 0, cv2.CV_32F).ravel()
    # warp, 1 and 3 channels, incl. partially and fully outside taps
    src1 = rng.random((61, 83), dtype=np.float32)
    src3 = rng.random((61, 83, 3), dtype=np.float32)
    Ms = np.array([[[0.99999, -0.0031, 5.3], [0.0031, 0.99999, -7.1]],
                   [[1.0, 0.0, 0.5], [0.0, 1.0, 0.5]],
                   [[0.9, 0.2, -20.25], [-0.2, 0.9, 30.75]]], np.float32)
    out.update(warp_src1=src1, warp_src3=src3, warp_M=Ms)
    out["warp_dst1"] = np.stack([cv2.warpAffine(src1, M, (83, 61), flags=cv2.INTER_LINEAR + cv2.WARP_INVERSE_MAP) for M in Ms])
    out["warp_dst3"] = np.stack([cv2.warpAffine(src3, M, (83, 61), flags=cv2.INTER_LINEAR + cv2.WARP_INVERSE_MAP) for M in Ms])
    # calibration arithmetic
    a = rng.random((16, 16, 3), dtype=np.float32)
    b = rng.random((16, 16, 3), dtype=np.float32) * 0.01
    d = rng.random((16, 16, 3), dtype=np.float32) * 0.01
    fl = rng.random((16, 16, 3), dtype=np.float32) + 0.5
    r = cv2.subtract(a, b, dtype=cv2.CV_32F)
    r = cv2.subtract(r, d, dtype=cv2.CV_32F)
    r = cv2.divide(r, fl, scale=1.0, dtype=cv2.CV_32F)
    out.update(cal_in=a, cal_bias=b, cal_dark=d, cal_flat=fl, cal_out=r)
    np.savez_compressed(os.path.join(HERE, "golden_cv2.npz"), **out)
    print("golden_cv2.npz:", {k: v.shape for k, v in out.items()})


def make_ref():
    assert oracle.ref() is not None, "oracle/_ref not built (needs /root/reference)"
    rng = np.random.default_rng(11)
    out = {}
    # HFTI: 64 systems
    A, B, MM, AO, BO, KR, IP = [], [], [], [], [], [], []
    for t in range(64):
        m = int(rng.integers(6, 121))
        a = np.zeros((3, 120), np.float32)
        b = np.zeros((2, 120), np.float32)
        x = rng.integers(0, 9576, m).astype(np.float32)
        y = rng.integers(0, 6388, m).astype(np.float32)
        if t % 3 == 0:
            x, y = y * 0.3, x * 1.0
        a[0, :m], a[1, :m], a[2, :m] = x, y, 1
        th = rng.uniform(-0.01, 0.01)
        b[0, :m] = np.round(np.cos(th) * x - np.sin(th) * y + rng.uniform(-10, 10) + rng.normal(0, 0.5, m))
        b[1, :m] = np.round(np.sin(th) * x + np.cos(th) * y + rng.uniform(-10, 10) + rng.normal(0, 0.5, m))
        ao, bo, kr, ip = oracle.hfti(a, b, m, which="ref")
        A.append(a); B.append(b); MM.append(m); AO.append(ao); BO.append(bo); KR.append(kr); IP.append(ip)
    out.update(hfti_a=np.stack(A), hfti_b=np.stack(B), hfti_m=np.array(MM, np.int32), hfti_a_out=np.stack(AO),
               hfti_b_out=np.stack(BO), hfti_krank=np.array(KR, np.int32), hfti_ip=np.array(IP, np.int32))
    # quicksort: 24 key vectors, padded
    keys = np.zeros((24, 600), np.float32)
    lens = np.zeros(24, np.int32)
    skeys = np.zeros((24, 600), np.float32)
    sids = np.zeros((24, 600), np.int32)
    for t in range(24):
        n = int(rng.integers(0, 600))
        x = rng.uniform(0.5, 0.9, n).astype(np.float32)
        if t % 4 == 0 and n > 10:
            x[rng.integers(0, n, n // 3)] = x[0]
        xs, ids = oracle.sort_keys(x, which="ref")
        keys[t, :n], lens[t], skeys[t, :n], sids[t, :n] = x, n, xs, ids
    out.update(sort_keys=keys, sort_len=lens, sort_out=skeys, sort_ids=sids)
    # alignment: 48 star-list pairs
    L1 = np.zeros((48, 64, 2), np.int32); L2 = np.zeros((48, 64, 2), np.int32)
    N1 = np.zeros(48, np.int32); N2 = np.zeros(48, np.int32)
    ERR = np.zeros(48, np.int32); K = np.zeros(48, np.int32)
    MT = np.zeros((48, 3, 120), np.int32); MX = np.zeros((48, 2, 3), np.float32)
    NT = np.zeros(48, np.int32)
    for t in range(48):
        xy1, xy2 = star_lists(rng, t)
        r = oracle.align(xy1, xy2, which="ref")
        idx, txy = oracle.triangles(xy1, which="ref")
        L1[t, :len(xy1)], L2[t, :len(xy2)], N1[t], N2[t] = xy1, xy2, len(xy1), len(xy2)
        ERR[t], K[t], MT[t], MX[t], NT[t] = r["err"], r["k"], r["matches"], r["M"], len(idx)
        if t == 0:
            out.update(tri_idx=idx, tri_xy=txy)
    out.update(al_xy1=L1, al_xy2=L2, al_n1=N1, al_n2=N2, al_err=ERR, al_k=K, al_matches=MT, al_M=MX, al_ntri=NT)
    # detection: 3 blended-blob images x 3 thresholds
    imgs = np.stack([blob_image(160, 120, 24, s) for s in range(3)])
    out["det_img"] = imgs
    ths = np.array([0.004, 0.02, 0.1], np.float32)
    out["det_th"] = ths
    for i in range(3):
        for j, th in enumerate(ths):
            st, nc = oracle.detect(imgs[i], float(th), float(np.float32(th * 4 / 3)), which="ref")
            out[f"det_{i}_{j}"] = st
            out[f"det_nc_{i}_{j}"] = np.int32(nc)
    np.savez_compressed(os.path.join(HERE, "golden_ref.npz"), **out)
    print("golden_ref.npz written:", len(out), "arrays")


if __name__ == "__main__":
    make_cv2()
    make_ref()
