"""Seeded synthetic frames (SURVEY.md §8(d) "Synthetic frame spec").

Host-side numpy generator used by the tests, smoke() and the bench's parity
samples: a fixed star field (Gaussian PSFs at sub-pixel centres, no saturation)
seen through a small rigid motion per frame, plus sensor-like dark / bias / flat
frames so that calibration is meaningful.  Sub-pixel motion + noise on purpose:
exact integer shifts make the reference's clipping loop reject every pair
(focas.cpp:338-369).
"""
import numpy as np


class Field:
    """A star field in reference-frame coordinates."""

    def __init__(self, width, height, nstars, seed=1234):
        rng = np.random.default_rng(seed)
        self.width, self.height = width, height
        self.x = rng.uniform(0, width, nstars)
        self.y = rng.uniform(0, height, nstars)
        self.amp = np.exp(rng.uniform(np.log(0.01), np.log(0.9), nstars))
        self.sigma = rng.uniform(1.2, 2.5, nstars)


def motion(index, width, height, max_shift=8.0, max_rot=0.003):
    """Rigid motion of frame `index` about the image centre: (angle, dx, dy). Frame 0 = identity."""
    if index == 0:
        return 0.0, 0.0, 0.0
    rng = np.random.default_rng(2000 + index)
    dx, dy = rng.uniform(-max_shift, max_shift, 2)
    return float(rng.uniform(-max_rot, max_rot)), float(dx), float(dy)


def render_sky(field, index, background=0.05):
    """Noise-free float64 sky of frame `index` (values in [0,1] units)."""
    w, h = field.width, field.height
    ang, dx, dy = motion(index, w, h)
    cx, cy = 0.5 * w, 0.5 * h
    ca, sa = np.cos(ang), np.sin(ang)
    xs = cx + ca * (field.x - cx) - sa * (field.y - cy) + dx
    ys = cy + sa * (field.x - cx) + ca * (field.y - cy) + dy
    img = np.full((h, w), background, np.float64)
    for x, y, a, s in zip(xs, ys, field.amp, field.sigma):
        r = int(np.ceil(6 * s))
        x0, x1 = max(int(x) - r, 0), min(int(x) + r + 1, w)
        y0, y1 = max(int(y) - r, 0), min(int(y) + r + 1, h)
        if x0 >= x1 or y0 >= y1:
            continue
        gx = np.exp(-0.5 * ((np.arange(x0, x1) + 0.5 - x) / s) ** 2)
        gy = np.exp(-0.5 * ((np.arange(y0, y1) + 0.5 - y) / s) ** 2)
        img[y0:y1, x0:x1] += a * np.outer(gy, gx)
    return img


def flat_true(width, height):
    yy, xx = np.mgrid[0:height, 0:width]
    r2 = ((xx - 0.5 * width) ** 2 + (yy - 0.5 * height) ** 2) / (0.25 * (width ** 2 + height ** 2))
    return 0.5 * (1.0 - 0.3 * r2)


def _to_u16(adu, channels, gains=(1.0, 0.9, 0.8)):
    if channels == 1:
        return np.clip(np.rint(adu), 0, 65535).astype(np.uint16)
    planes = [np.clip(np.rint(adu * g), 0, 65535).astype(np.uint16) for g in gains[:channels]]
    return np.ascontiguousarray(np.stack(planes, axis=2))


def light(field, index, channels=3, dark_adu=0.0, bias_adu=0.0, vignette=False, noise=0.002):
    """u16 HWC light frame `index` (0 = the reference)."""
    sky = render_sky(field, index)
    if vignette:
        sky = sky * (flat_true(field.width, field.height) / 0.5)
    rng = np.random.default_rng(1000 + index)
    adu = (sky + rng.normal(0.0, noise, sky.shape)) * 65535.0 + dark_adu + bias_adu
    return _to_u16(adu, channels)


def dark(width, height, index, channels=3, level=300.0, bias_adu=0.0):
    rng = np.random.default_rng(3000 + index)
    return _to_u16(level + bias_adu + rng.normal(0.0, 20.0, (height, width)), channels, (1.0, 1.0, 1.0))


def bias(width, height, index, channels=3, level=100.0):
    rng = np.random.default_rng(4000 + index)
    return _to_u16(level + rng.normal(0.0, 5.0, (height, width)), channels, (1.0, 1.0, 1.0))


def flat(width, height, index, channels=3, bias_adu=0.0):
    rng = np.random.default_rng(5000 + index)
    adu = flat_true(width, height) * 65535.0 + bias_adu + rng.normal(0.0, 50.0, (height, width))
    return _to_u16(np.maximum(adu, 1.0), channels, (1.0, 1.0, 1.0))


def dataset(width, height, nlights, channels=3, nstars=None, ndarks=0, nflats=0, nbias=0, ndarkflats=0, seed=1234):
    """A whole seeded stack: dict(ref, lights, darks, flats, bias, darkflats) of u16 arrays."""
    if nstars is None:
        nstars = max(60, int(600 * (width * height) / 24e6))
    f = Field(width, height, nstars, seed)
    b = 100.0 if nbias else 0.0
    d = 300.0 if ndarks else 0.0
    vign = nflats > 0
    frames = [light(f, i, channels, d, b, vign) for i in range(nlights + 1)]
    return dict(
        field=f,
        ref=frames[0], lights=frames[1:],
        darks=[dark(width, height, i, channels, 300.0, b) for i in range(ndarks)],
        bias=[bias(width, height, i, channels) for i in range(nbias)],
        flats=[flat(width, height, i, channels, b) for i in range(nflats)],
        darkflats=[bias(width, height, 100 + i, channels, 100.0 + 5.0) if nbias else
                   bias(width, height, 100 + i, channels, 5.0) for i in range(ndarkflats)],
    )
