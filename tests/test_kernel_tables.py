"""Constant tables inside the CUDA sources that can be proven correct on the CPU."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_median_network():
    """kMed25 (k_sky.cuh) must leave the median of 25 in wire 12: exhaustive 0/1-principle check, 2^25 inputs bit-parallel.
    medianBlur(5) is an exact selection, so any correct selection network is bit-exact."""
    src = open(os.path.join(ROOT, "openskystacker_b200", "csrc", "k_sky.cuh")).read()
    body = re.search(r"kMed25\[198\]\s*=\s*\{([^}]*)\}", src).group(1)
    v = [int(t) for t in re.findall(r"\d+", body)]
    assert len(v) == 198
    pairs = list(zip(v[0::2], v[1::2]))
    words = (1 << 25) // 64
    idx = np.arange(words, dtype=np.uint64)
    wires = []
    for i in range(25):
        if i < 6:
            pat = sum(((n >> i) & 1) << n for n in range(64))
            wires.append(np.full(words, pat, dtype=np.uint64))
        else:
            wires.append(np.where(((idx >> np.uint64(i - 6)) & np.uint64(1)) == 1, np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64(0)))
    bits = np.arange(64, dtype=np.uint64)
    ones = np.zeros((words, 64), dtype=np.uint8)
    for w in wires:
        ones += ((w[:, None] >> bits[None, :]) & np.uint64(1)).astype(np.uint8)
    for a, b in pairs:
        assert a < b
        wires[a], wires[b] = wires[a] & wires[b], wires[a] | wires[b]
    got = ((wires[12][:, None] >> bits[None, :]) & np.uint64(1)).astype(bool)
    assert np.array_equal(got, ones >= 13)


def test_focas_vote_thresholds_are_the_floats_just_above_the_double_constants():
    """k_focas compares float differences with float constants instead of converting every window entry to double:
    for a float d, (double)d < t  <=>  d < f where f is the smallest float32 >= t (t itself is not a float32)."""
    import re
    src = open(os.path.join(os.path.dirname(__file__), "..", "openskystacker_b200", "csrc", "k_focas.cuh")).read()
    bits = [int(x, 16) for x in re.search(r"kTolUp = __uint_as_float\((0x[0-9a-f]+)u\), kTol2Up = __uint_as_float\((0x[0-9a-f]+)u\)", src).groups()]
    for b, t in zip(bits, (0.002, 0.002 * 0.002)):
        f = np.array([b], np.uint32).view(np.float32)[0]
        below = np.nextafter(f, np.float32(-np.inf))
        assert float(below) < t <= float(f)
        # exhaustive around the boundary: the float test and the double test agree
        for d in (below, f, np.nextafter(below, np.float32(-np.inf)), np.nextafter(f, np.float32(np.inf)), np.float32(-t), np.float32(0)):
            assert (float(d) < t) == bool(d < f)
