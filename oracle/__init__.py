"""TEST INFRASTRUCTURE — ctypes doors to the CPU oracle.

`liboss_oracle.so`  = oracle/oss_oracle.c, the plain-C restatement of libstacker's
per-frame pipeline.  `_ref/liboss_ref.so` = the reference's own focas.cpp /
adjoiningpixel.cpp / HFTI object code behind oracle/ref_harness.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (openskystacker_b200/)
never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
NOBJS, MAX_MATCH, MAX_TRI = 40, 120, 9880
LD = MAX_MATCH + 1


class Star(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("area", C.c_int),
                ("peak", C.c_float), ("value", C.c_float)]


class DetectInfo(C.Structure):
    _fields_ = [("mean", C.c_double), ("sigma", C.c_double), ("threshold", C.c_float),
                ("minpeak", C.c_float), ("ncomponents", C.c_int), ("nstars", C.c_int)]


class FrameReport(C.Structure):
    _fields_ = [("nstars", C.c_int), ("k", C.c_int), ("err", C.c_int),
                ("M", C.c_float * 6), ("threshold", C.c_float)]


STAR_DTYPE = np.dtype([("x", "i4"), ("y", "i4"), ("area", "i4"), ("peak", "f4"), ("value", "f4")])


def build(force=False):
    """Compile the checker (building it is not using it)."""
    so = os.path.join(_HERE, "liboss_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("oss_oracle.c", "oss_oracle.h")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src)
    if stale or (os.path.isdir("/root/reference/libstacker/src")
                 and not os.path.exists(os.path.join(_HERE, "_ref", "liboss_ref.so"))):
        subprocess.run(["make", "-C", _HERE, "--no-print-directory"], check=True,
                       stdout=subprocess.DEVNULL)


_lib = None
_ref = None


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(os.path.join(_HERE, "liboss_oracle.so"))
        _lib.orc_roi_stats.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        for name in ("orc_detect", "orc_deblend", "orc_get_stars", "orc_triangles", "orc_find_transform",
                     "orc_align", "orc_process_light"):
            getattr(_lib, name).restype = C.c_int
    return _lib


def ref():
    """The reference's own object code, or None when oracle/_ref was never built."""
    global _ref
    if _ref is None:
        path = os.path.join(_HERE, "_ref", "liboss_ref.so")
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            return None
        _ref = C.CDLL(path)
        for name in ("ref_detect", "ref_deblend", "ref_triangles", "ref_align"):
            getattr(_ref, name).restype = C.c_int
    return _ref


# --------------------------------------------------------------------------- primitives
def u16_to_f32(a):
    a = np.ascontiguousarray(a, dtype=np.uint16)
    out = np.empty(a.shape, np.float32)
    lib().orc_u16_to_f32(C.c_void_p(a.ctypes.data), C.c_void_p(out.ctypes.data), C.c_size_t(a.size))
    return out


def stack_mean(frames, sub1=None, sub2=None):
    frames = [_f32(f) for f in frames]
    n = len(frames)
    ptrs = (C.c_void_p * n)(*[f.ctypes.data for f in frames])
    out = np.empty(frames[0].shape, np.float32)
    s1 = _f32(sub1) if sub1 is not None else None
    s2 = _f32(sub2) if sub2 is not None else None
    lib().orc_stack_mean(ptrs, C.c_int(n), C.c_size_t(out.size),
                         C.c_void_p(s1.ctypes.data if s1 is not None else None),
                         C.c_void_p(s2.ctypes.data if s2 is not None else None),
                         C.c_void_p(out.ctypes.data))
    return out


def flat_normalize(flat):
    flat = _f32(flat).copy()
    ch = 1 if flat.ndim == 2 else flat.shape[2]
    lib().orc_flat_normalize(C.c_void_p(flat.ctypes.data), C.c_size_t(flat.size // ch), C.c_int(ch))
    return flat


def calibrate(img, bias=None, dark=None, flat=None):
    img = _f32(img).copy()
    ms = [(_f32(m) if m is not None else None) for m in (bias, dark, flat)]
    lib().orc_calibrate(C.c_void_p(img.ctypes.data), C.c_size_t(img.size),
                        *[C.c_void_p(m.ctypes.data if m is not None else None) for m in ms])
    return img


def gray(img):
    img = _f32(img)
    ch = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty(img.shape[:2], np.float32)
    lib().orc_gray(C.c_void_p(img.ctypes.data), C.c_int(ch), C.c_void_p(out.ctypes.data), C.c_size_t(out.size))
    return out


def resize_linear(src, dw, dh):
    src = _f32(src)
    out = np.empty((dh, dw), np.float32)
    lib().orc_resize_linear(C.c_void_p(src.ctypes.data), C.c_int(src.shape[1]), C.c_int(src.shape[0]),
                            C.c_void_p(out.ctypes.data), C.c_int(dw), C.c_int(dh))
    return out


def _unary(fn, src):
    src = _f32(src)
    out = np.empty_like(src)
    fn(C.c_void_p(src.ctypes.data), C.c_void_p(out.ctypes.data), C.c_int(src.shape[1]), C.c_int(src.shape[0]))
    return out


def median5(src):
    return _unary(lib().orc_median5, src)


def gauss31(src):
    return _unary(lib().orc_gauss31, src)


def sky_background(g):
    return _unary(lib().orc_sky_background, g)


def roi_stats(img):
    img = _f32(img)
    m, s = C.c_double(), C.c_double()
    lib().orc_roi_stats(C.c_void_p(img.ctypes.data), C.c_int(img.shape[1]), C.c_int(img.shape[0]),
                        C.byref(m), C.byref(s))
    return m.value, s.value


def _stars_out(buf, n):
    return np.frombuffer(buf, dtype=STAR_DTYPE, count=min(n, len(buf)))[:n].copy()


def detect(stars_img, threshold, minpeak, cap=1 << 18, which="oracle"):
    """Components + deblend + centroids on a sky-subtracted image (copied, not destroyed)."""
    img = _f32(stars_img).copy()
    buf = (Star * cap)()
    nc = C.c_int()
    fn = lib().orc_detect if which == "oracle" else ref().ref_detect
    n = fn(C.c_void_p(img.ctypes.data), C.c_int(img.shape[1]), C.c_int(img.shape[0]),
           C.c_float(threshold), C.c_float(minpeak), buf, C.c_int(cap), C.byref(nc))
    return _stars_out(buf, n), nc.value


def deblend(px, py, pv, step, cap=4096, which="oracle"):
    px = np.ascontiguousarray(px, np.int32)
    py = np.ascontiguousarray(py, np.int32)
    pv = _f32(pv)
    buf = (Star * cap)()
    fn = lib().orc_deblend if which == "oracle" else ref().ref_deblend
    n = fn(_p(px, C.c_int), _p(py, C.c_int), _p(pv, C.c_float), C.c_int(len(px)), C.c_float(step), buf, C.c_int(cap))
    return _stars_out(buf, n)


def get_stars(img, threshold_coeff, cap=1 << 18):
    img = _f32(img)
    ch = 1 if img.ndim == 2 else img.shape[2]
    buf = (Star * cap)()
    info = DetectInfo()
    n = lib().orc_get_stars(C.c_void_p(img.ctypes.data), C.c_int(img.shape[1]), C.c_int(img.shape[0]), C.c_int(ch),
                            C.c_int(threshold_coeff), buf, C.c_int(cap), C.byref(info))
    return _stars_out(buf, n), info


def _xy(stars_or_xy):
    a = stars_or_xy
    if isinstance(a, np.ndarray) and a.dtype.names:
        a = np.stack([a["x"], a["y"]], axis=1)
    return np.ascontiguousarray(a, np.int32).reshape(-1, 2)


def triangles(xy, which="oracle"):
    xy = _xy(xy)
    idx = np.zeros((MAX_TRI, 3), np.int32)
    txy = np.zeros((MAX_TRI, 2), np.float32)
    if which == "oracle":
        n = lib().orc_triangles(_p(xy, C.c_int), C.c_int(len(xy)), _p(idx, C.c_int), _p(txy, C.c_float))
    else:
        n = ref().ref_triangles(_p(xy, C.c_int), C.c_int(len(xy)), _p(idx, C.c_int), _p(txy, C.c_float), C.c_int(MAX_TRI))
    return idx[:n].copy(), txy[:n].copy()


def sort_keys(x, which="oracle"):
    x = _f32(x).copy()
    ids = np.arange(len(x), dtype=np.int32)
    fn = lib().orc_sort_keys if which == "oracle" else ref().ref_sort_keys
    fn(_p(x, C.c_float), _p(ids, C.c_int), C.c_int(len(x)))
    return x, ids


def align(xy1, xy2, which="oracle"):
    """-> dict(err, k, matches[3,120], M[2,3], table[40,40] (oracle only))."""
    xy1, xy2 = _xy(xy1), _xy(xy2)
    k = C.c_int()
    M = np.zeros(6, np.float32)
    if which == "oracle":
        matches = np.zeros((3, LD), np.int32)
        table = np.zeros((NOBJS, NOBJS), np.int32)
        err = lib().orc_align(_p(xy1, C.c_int), C.c_int(len(xy1)), _p(xy2, C.c_int), C.c_int(len(xy2)),
                              C.byref(k), _p(matches, C.c_int), _p(table, C.c_int), _p(M, C.c_float))
        return dict(err=err, k=k.value, matches=matches[:, :MAX_MATCH].copy(), M=M.reshape(2, 3), table=table)
    matches = np.zeros((3, MAX_MATCH), np.int32)
    err = ref().ref_align(_p(xy1, C.c_int), C.c_int(len(xy1)), _p(xy2, C.c_int), C.c_int(len(xy2)),
                          C.byref(k), _p(matches, C.c_int), _p(M, C.c_float))
    return dict(err=err, k=k.value, matches=matches, M=M.reshape(2, 3), table=None)


def hfti(a, b, m, tau=0.1, which="oracle"):
    """a: (3,120) f32 (column-major m x 3), b: (2,120).  Returns (a, b, krank, ip) after the call."""
    a = _f32(a).copy()
    b = _f32(b).copy()
    krank = C.c_int()
    rnorm = (C.c_float * 2)()
    ip = (C.c_int * 3)()
    if which == "oracle":
        h = (C.c_float * 3)()
        g = (C.c_float * 3)()
        lib().orc_hfti(_p(a, C.c_float), C.c_int(MAX_MATCH), C.c_int(m), C.c_int(3), _p(b, C.c_float),
                       C.c_int(MAX_MATCH), C.c_int(2), C.c_float(tau), C.byref(krank), rnorm, h, g, ip)
    else:
        ref().ref_hfti(_p(a, C.c_float), C.c_int(m), _p(b, C.c_float), C.c_float(tau), C.byref(krank), rnorm, ip)
    return a, b, krank.value, list(ip)


def warp_affine(src, M):
    src = _f32(src)
    ch = 1 if src.ndim == 2 else src.shape[2]
    M = _f32(M).reshape(6)
    out = np.empty_like(src)
    lib().orc_warp_affine(C.c_void_p(src.ctypes.data), C.c_void_p(out.ctypes.data), C.c_int(src.shape[1]),
                          C.c_int(src.shape[0]), C.c_int(ch), _p(M, C.c_float))
    return out


def process_light(raw, bias, dark, flat, threshold_coeff, ref_cal, ref_xy, total):
    """One pass of processConcurrent's loop body; `total` (f32, HWC) is accumulated in place."""
    raw = np.ascontiguousarray(raw, np.uint16)
    h, w = raw.shape[:2]
    ch = 1 if raw.ndim == 2 else raw.shape[2]
    rep = FrameReport()
    ms = [(m.ctypes.data if m is not None else None) for m in (bias, dark, flat)]
    if ref_xy is not None:
        ref_xy = _xy(ref_xy)
    assert total.dtype == np.float32 and total.flags.c_contiguous
    err = lib().orc_process_light(C.c_void_p(raw.ctypes.data), C.c_int(w), C.c_int(h), C.c_int(ch),
                                  C.c_void_p(ms[0]), C.c_void_p(ms[1]), C.c_void_p(ms[2]), C.c_int(threshold_coeff),
                                  C.c_void_p(ref_cal.ctypes.data if ref_cal is not None else None),
                                  _p(ref_xy, C.c_int) if ref_xy is not None else None,
                                  C.c_int(len(ref_xy) if ref_xy is not None else 0),
                                  C.c_void_p(total.ctypes.data), C.byref(rep))
    return err, rep


def scale(img, total_valid):
    img = _f32(img).copy()
    lib().orc_scale(C.c_void_p(img.ctypes.data), C.c_size_t(img.size), C.c_int(total_valid))
    return img


def stack(ref_raw, lights, threshold_coeff, bias_frames=None, dark_frames=None, darkflat_frames=None,
          flat_frames=None, redetect_ref=False):
    """ImageStackerPrivate::process (imagestacker.cpp:262-361) on in-memory u16 frames.

    Returns dict(image, total_valid, reports, masters, ref_stars)."""
    conv = u16_to_f32
    bias = stack_mean([conv(f) for f in bias_frames]) if bias_frames else None
    dark = stack_mean([conv(f) for f in dark_frames], bias) if dark_frames else None
    darkflat = stack_mean([conv(f) for f in darkflat_frames], bias) if darkflat_frames else None
    flat = flat_normalize(stack_mean([conv(f) for f in flat_frames], bias, darkflat)) if flat_frames else None
    ref_cal = calibrate(conv(ref_raw), bias, dark, flat)
    ref_stars, ref_info = get_stars(ref_cal, threshold_coeff)
    total = np.zeros_like(ref_cal)
    reports, valid = [], 1
    for raw in lights:
        err, rep = process_light(raw, bias, dark, flat, threshold_coeff, ref_cal,
                                 None if redetect_ref else ref_stars, total)
        reports.append(rep)
        valid += 0 if err else 1
    # workingImage = ref.clone(); workingImage += partial (imagestacker.cpp:287,349)
    image = ref_cal + total
    if valid >= 2:
        image = scale(image, valid)
    return dict(image=image, total_valid=valid, reports=reports, ref_stars=ref_stars, ref_info=ref_info,
                masters=dict(bias=bias, dark=dark, darkflat=darkflat, flat=flat), ref_cal=ref_cal)
