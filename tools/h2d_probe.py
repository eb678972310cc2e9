"""Dev aid (GPU box): what the PCIe link gives for pinned host -> device copies of frame-sized buffers, with 1, 2 and 3 copy
streams (the e2e number of bench.py is bound by this)."""
import time

import torch

n = 144_000_000  # one 24 MP RGB16 frame
host = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(6)]
dev = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(6)]
for ns in (1, 2, 3):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for it in range(4):
            for i in range(6):
                with torch.cuda.stream(streams[i % ns]):
                    dev[i].copy_(host[i], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"{ns} stream(s): {24 * n / dt / 1e9:.2f} GB/s")
# both directions at once
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
back = torch.empty(n, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize()
t0 = time.perf_counter()
for it in range(4):
    for i in range(6):
        with torch.cuda.stream(s1):
            dev[i].copy_(host[i], non_blocking=True)
    with torch.cuda.stream(s2):
        back.copy_(dev[0], non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"H2D beside a D2H stream: {24 * n / dt / 1e9:.2f} GB/s H2D")
