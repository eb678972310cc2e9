"""Dev aid: host-side duration of each C-ABI call of one bench step (device-resident inputs)."""
import ctypes as Ct, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openskystacker_b200 as ob
from openskystacker_b200 import binding as B
import bench

W, H, C = 6000, 4000, 3
eng = ob.Engine(0)
eng.set_pipelines(3)
eng.set_geometry(W, H, C, np.uint16, 17)
inp = bench.Inputs(eng, W, H, 50, 10)
arr = lambda ps: ((Ct.c_void_p * len(ps))(*ps), len(ps))
cal = (arr(inp.bias), arr(inp.darks), arr(inp.flats))
lights, nl = arr(inp.lights)
L, h = eng._L, eng._h
for it in range(4):
    eng.sync()
    t = [time.perf_counter()]
    for kind, (a, n) in zip((B.BIAS, B.DARK, B.FLAT), cal):
        L.oss_build_master(h, kind, a, n, B.DEVICE); t.append(time.perf_counter())
    L.oss_set_reference(h, Ct.c_void_p(inp.ref), B.DEVICE, 20); t.append(time.perf_counter())
    L.oss_add_lights(h, lights, nl, B.DEVICE); t.append(time.perf_counter())
    tv = Ct.c_int(); L.oss_finish(h, None, Ct.byref(tv)); t.append(time.perf_counter())
    d = [round(1e3 * (b - a), 3) for a, b in zip(t, t[1:])]
    print("ms: bias %s dark %s flat %s set_reference %s add_lights %s finish %s total %s" % (*d, round(1e3 * (t[-1] - t[0]), 3)))
