"""Dev aid (multi-GPU box, under torchrun): what the box gives when EVERY rank copies pinned host frames to its GPU at the same
time (the N-GPU e2e number of bench.py is bound by this), with and without the rank's CPU affinity set to the GPU's NUMA node
before the pinned allocation.  Rank 0 also prints the topology the box shows."""
import os
import subprocess
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("gloo")


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=30).stdout
    except Exception as e:  # noqa: BLE001
        return f"({e})"


if rank == 0:
    print(sh("nvidia-smi topo -m"))
    print(sh("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"))
    print(sh("for d in /sys/bus/pci/devices/*; do c=$(cat $d/class); if [ \"$c\" = 0x030200 ]; then echo $d numa $(cat $d/numa_node) cpus $(cat $d/local_cpulist); fi; done"))
    print(sh("free -g | head -2"), flush=True)


def affinity_from_nvml():
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
    return cpus


n = 144_000_000
NB = 6


def measure(tag):
    host = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(NB)]
    for h in host:
        h.fill_(1)
    dev = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(NB)]
    s = torch.cuda.Stream()
    out = []
    for rep in range(3):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for it in range(6):
            for i in range(NB):
                with torch.cuda.stream(s):
                    dev[i].copy_(host[i], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out.append(6 * NB * n / dt / 1e9)
    g = [None] * world
    dist.all_gather_object(g, (rank, round(max(out), 1), sorted(os.sched_getaffinity(0))[:1], len(os.sched_getaffinity(0))))
    if rank == 0:
        print(tag, "per-rank GB/s:", [x[1] for x in g], "sum", round(sum(x[1] for x in g), 1), "affinity(first cpu, count):", [(x[2], x[3]) for x in g], flush=True)
    del host, dev


measure("default      ")
try:
    cpus = affinity_from_nvml()
    if cpus:
        os.sched_setaffinity(0, cpus)
    measure("nvml affinity")
except Exception as e:  # noqa: BLE001
    if rank == 0:
        print("nvml affinity unavailable:", e)
# one rank alone (rank 0), the others idle
if world > 1:
    os.sched_setaffinity(0, range(os.cpu_count()))
    for solo in (0, world - 1):
        dist.barrier()
        if rank == solo:
            host = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(NB)]
            dev = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(NB)]
            for rep in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for it in range(4):
                    for i in range(NB):
                        dev[i].copy_(host[i], non_blocking=True)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            print(f"rank {solo} alone: {4 * NB * n / dt / 1e9:.1f} GB/s", flush=True)
            del host, dev
        dist.barrier()
dist.destroy_process_group()
