"""Dev aid: time oss_detect_sweep (config 5: -s threshold sweep 1..100 on a 61 MP mono frame) and the batched detector."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openskystacker_b200 as ob
from openskystacker_b200 import synth
import bench

W, H = 9576, 6388
eng = ob.Engine(0)
eng.set_geometry(W, H, 1, np.uint16, int(os.environ.get("SLOTS", "16")))
field = synth.Field(W, H, 4000, 1234)
p = eng.device_alloc(eng.frame_bytes())
eng.synth_light(p, bench.star_motion(field, 0, W, H), pedestal_adu=400.0, vignette=True, seed=1000)
frame = eng.from_device(p, (H, W), np.uint16)
for it in range(3):
    t0 = time.perf_counter()
    counts = eng.detect_sweep(frame, 1, 100)
    t1 = time.perf_counter()
    print("sweep 1..100: %.1f ms, counts[0,4,19,49,99] = %s" % (1e3 * (t1 - t0), [counts[i] for i in (0, 4, 19, 49, 99)]))
n = 40
frames = [frame] * n
for it in range(3):
    t0 = time.perf_counter()
    counts = eng.detect_stars_batch(frames, 5)
    t1 = time.perf_counter()
    print("detect_stars_batch: %d x 61 MP host frames at t=5: %.1f ms = %.0f frames/s (%.1f GB/s H2D), stars %d" %
          (n, 1e3 * (t1 - t0), n / (t1 - t0), n * frame.nbytes / (t1 - t0) / 1e9, counts[0]))
import ctypes as Ct
L = eng._L
L.oss_host_alloc.restype = Ct.c_void_p
pinned = [L.oss_host_alloc(Ct.c_size_t(frame.nbytes)) for _ in range(n)]
for q in pinned:
    Ct.memmove(q, frame.ctypes.data, frame.nbytes)
arr = (Ct.c_void_p * n)(*pinned)
counts = np.zeros(n, np.int32)
for it in range(3):
    t0 = time.perf_counter()
    st = L.oss_detect_stars_batch(eng._h, arr, n, 0, 5, counts.ctypes.data_as(Ct.c_void_p))
    t1 = time.perf_counter()
    print("pinned: status %d, %.1f ms = %.0f frames/s (%.1f GB/s H2D), stars %d" % (st, 1e3 * (t1 - t0), n / (t1 - t0), n * frame.nbytes / (t1 - t0) / 1e9, counts[0]))
