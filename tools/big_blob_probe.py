"""Dev aid (GPU box): detection time and parity on frames with ONE very large component (a saturated star / galaxy core):
components above 256 pixels go through k_components_big (one CTA per component; before it: one thread, 6 - 37 ms for these
frames).  Second part: the large-component test image (ramped noisy lattices of 260 ... 8 500 px, ~100 stars out of ~10
components) in a batch of 8 frames."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import openskystacker_b200 as ob   # noqa: E402
import oracle                      # noqa: E402

w, h = 3000, 2000
eng = ob.Engine(0)
eng.set_geometry(w, h, 1, np.float32, 2)
yy, xx = np.mgrid[0:h, 0:w]
rng = np.random.default_rng(5)
for radius in (9, 18, 30, 45, 55, 60):
    img = (0.05 + rng.normal(0, 0.002, (h, w))).astype(np.float32)
    r2 = (xx - 1500.3) ** 2 + (yy - 1000.7) ** 2
    img += (0.3 * np.exp(-r2 / (2 * (radius / 2.0) ** 2)) + 0.5 * (r2 < radius ** 2)).astype(np.float32)
    eng.detect_stars(img, 5)
    t0 = time.perf_counter()
    stars, info = eng.detect_stars(img, 5)
    t1 = time.perf_counter()
    o_stars, o_info = oracle.get_stars(img, 5)
    t2 = time.perf_counter()
    same = len(stars) == len(o_stars) and all((a["x"], a["y"], a["area"]) == (b["x"], b["y"], b["area"]) for a, b in zip(stars, o_stars))
    big = max((int(s["area"]) for s in stars), default=0)
    print(f"radius {radius}: candidates {info.ncandidates}, stars {len(stars)} (largest area {big}), GPU {1e3 * (t1 - t0):.1f} ms, "
          f"oracle {1e3 * (t2 - t1):.1f} ms, identical {same}")
sys.path.insert(0, "tests")
from test_gpu_parity import large_component_image   # noqa: E402
w, h = 1600, 1000
eng.set_geometry(w, h, 1, np.float32, 8)
for seed, t in ((21, 1), (22, 2), (23, 3)):
    g = large_component_image(w, h, seed)
    eng.detect_stars_batch([g] * 8, t)
    t0 = time.perf_counter()
    counts = eng.detect_stars_batch([g] * 8, t)
    t1 = time.perf_counter()
    o_stars, _ = oracle.get_stars(g, t)
    t2 = time.perf_counter()
    print(f"lattices seed {seed} t {t}: stars {counts[0]} (oracle {len(o_stars)}), GPU {1e3 * (t1 - t0) / 8:.2f} ms per frame (batch of 8, "
          f"incl. H2D), oracle {1e3 * (t2 - t1):.1f} ms")
# a satellite trail: one component of ~14 000 px, the deblender bypass
from test_gpu_parity import blob_image   # noqa: E402
w, h = 2400, 1600
eng.set_geometry(w, h, 1, np.float32, 2)
r = np.random.default_rng(31)
g = blob_image(w, h, 30, 31).astype(np.float64)
yy, xx = np.mgrid[0:h, 0:w]
d = np.abs((yy - 100) - 0.55 * (xx - 50)) / np.hypot(1, 0.55)
g = (g + (d < 2.6) * (0.5 + 0.1 * np.sin(xx / 37.0) + r.normal(0, 0.02, (h, w)))).astype(np.float32)
eng.detect_stars(g, 5)
t0 = time.perf_counter()
stars, info = eng.detect_stars(g, 5)
t1 = time.perf_counter()
o_stars, _ = oracle.get_stars(g, 5)
t2 = time.perf_counter()
print(f"trail: stars {len(stars)} (oracle {len(o_stars)}), largest area {max(int(s['area']) for s in stars)}, GPU {1e3 * (t1 - t0):.1f} ms, "
      f"oracle {1e3 * (t2 - t1):.1f} ms")
eng.close()
