import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
os.environ["OSS_B200_FOCAS_CLK"] = "1"
import openskystacker_b200 as ob, oracle
e = ob.Engine(0); e.set_geometry(64, 64, 1, np.float32, 1)
rng = np.random.default_rng(5)
bad = 0
for t in range(300):
    n1 = int(rng.integers(3, 70)) if t % 3 else 60
    xy1 = np.stack([rng.integers(0, 6000, n1), rng.integers(0, 800, n1)], 1).astype(np.int32)
    th = rng.uniform(-0.003, 0.003); c, s = np.cos(th), np.sin(th); sh = rng.uniform(-8, 8, 2)
    xy2 = np.stack([c * xy1[:, 0] - s * xy1[:, 1] + sh[0], s * xy1[:, 0] + c * xy1[:, 1] + sh[1]], 1)
    xy2 = np.floor(xy2 + rng.normal(0, 0.4, (n1, 2))).astype(np.int32)
    if t == 5: os.environ.pop("OSS_B200_FOCAS_CLK")
    got, want = e.align_lists(xy1, xy2), oracle.align(xy1, xy2)
    ok = np.array_equal(got["table"], want["table"]) and got["k"] == want["k"] and got["err"] == want["err"] and np.array_equal(got["M"].view(np.uint32), want["M"].view(np.uint32))
    bad += not ok
print("mismatches", bad, "of 300")
