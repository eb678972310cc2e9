"""TEST / BASELINE INFRASTRUCTURE — the reference's CPU path driven the way the reference
itself runs it: OpenCV (cv2, optimised mode: IPP / AVX, the library libstacker links) for
every cv:: call, the reference's OWN object code (oracle/_ref: adjoiningpixel.cpp, focas.cpp,
Fortran HFTI) for deblending, triangles, voting and the fit, and the strided worker loop of
processConcurrent (libstacker/src/util.cpp:724-758).  Only the glue is restated.  Used by
bench.py's cpu_baseline / --impl reference legs as the fastest honest CPU baseline; parity
work uses the plain-C oracle instead (cv2's optimised kernels differ from its plain ones by
~1e-7)."""
import ctypes as C
from concurrent.futures import ThreadPoolExecutor

import numpy as np

import oracle

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def available():
    return cv2 is not None and oracle.ref() is not None


def get_stars(img, t):
    """StarDetectorImpl::getStars, stardetector.cpp:108-147."""
    gray = img if img.ndim == 2 else cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    h, w = gray.shape
    sky = cv2.resize(gray, (w // 8, h // 8))
    sky = cv2.medianBlur(sky, 5)
    sky = cv2.resize(sky, (w, h))
    sky = cv2.GaussianBlur(sky, (31, 31), 0)
    stars = gray - sky
    roi = stars[h // 20:h // 20 + int(h * 0.9), w // 20:w // 20 + int(w * 0.9)]
    _, sd = cv2.meanStdDev(roi)
    th = np.float32(sd[0, 0] * t * 1.5)
    mp = np.float32(sd[0, 0] * t * 2.0)
    out, _ = oracle.detect(stars, float(th), float(mp), which="ref")
    return out


def process_light(raw, masters, t, ref_cal):
    """getCalibratedImage + generateAlignedImage (util.cpp:266-274,557-582); returns (aligned or None)."""
    img = raw.astype(np.float32) * np.float32(1 / 65535.0)
    bias, dark, flat = masters
    if bias is not None:
        img = cv2.subtract(img, bias)
    if dark is not None:
        img = cv2.subtract(img, dark)
    if flat is not None:
        img = cv2.divide(img, flat)
    l1 = get_stars(ref_cal, t)  # the reference re-detects the ref stars for every light (util.cpp:559)
    l2 = get_stars(img, t)
    r = oracle.align(l1, l2, which="ref")
    if r["err"]:
        return None
    h, w = img.shape[:2]
    return cv2.warpAffine(img, r["M"], (w, h), flags=cv2.INTER_LINEAR + cv2.WARP_INVERSE_MAP)


def stack_lights(lights, masters, t, ref_cal, threads):
    """processConcurrent's strided fan-out (util.cpp:738) + the serial combine (imagestacker.cpp:348-351).
    Returns (sum image, valid count)."""
    cv2.setNumThreads(1)  # one OpenCV thread per worker, like QtConcurrent workers on a loaded box

    def worker(i):
        part = np.zeros_like(ref_cal)
        valid = 0
        for k in range(i, len(lights), threads):
            a = process_light(lights[k], masters, t, ref_cal)
            if a is not None:
                part += a
                valid += 1
        return part, valid

    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(worker, range(threads)))
    total = ref_cal.copy()
    valid = 1
    for p, v in parts:
        total += p
        valid += v
    return total, valid
