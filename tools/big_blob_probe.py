"""Dev aid (GPU box): detection time and parity on frames with ONE very large component (a saturated star / galaxy core):
components of 300 ... 9 900 pixels go through k_components' large-component path (deblending replayed by one thread)."""
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
eng.close()
