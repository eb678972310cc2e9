"""N > 1: frames shard across ranks, one reduce of (sum image, valid count) to rank 0.
CPU: two gloo ranks run the sharding + combine with the oracle as the per-rank worker and
must reproduce the single-process reference result.  GPU (needs >= 2 devices): two NCCL
ranks through the C ABI (oss_comm_*) against the single-GPU result and the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from openskystacker_b200 import synth
from openskystacker_b200.distributed import light_shard

W, H = 1000, 700


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def dataset(full_cal=False):
    if full_cal:
        return synth.dataset(W, H, 5, 1, nstars=120, ndarks=2, nflats=3, nbias=2, seed=1234)
    return synth.dataset(W, H, 5, 1, nstars=120, ndarks=2, seed=1234)


def test_light_shard_is_the_reference_stride():
    for n in (0, 1, 5, 50, 401):
        for world in (1, 2, 3, 8):
            shards = [light_shard(n, r, world) for r in range(world)]
            assert sorted(i for s in shards for i in s) == list(range(n))
            for r, s in enumerate(shards):
                assert s == list(range(r, n, world))          # util.cpp:738
    with pytest.raises(ValueError):
        light_shard(4, 2, 2)


def _gloo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds = dataset()
    conv = oracle.u16_to_f32
    # rank 0 builds the master + calibrates/detects the reference, then broadcasts (here: as arrays over gloo)
    if rank == 0:
        dark = oracle.stack_mean([conv(f) for f in ds["darks"]])
        ref_cal = oracle.calibrate(conv(ds["ref"]), None, dark, None)
        stars, _ = oracle.get_stars(ref_cal, 20)
        xy = np.stack([stars["x"], stars["y"]], 1).astype(np.int32)
        meta = torch.tensor([len(xy)])
    else:
        dark = np.zeros((H, W), np.float32)
        meta = torch.tensor([0])
    dist.broadcast(meta, 0)
    if rank != 0:
        xy = np.zeros((int(meta[0]), 2), np.int32)
    td, tx = torch.from_numpy(dark), torch.from_numpy(xy)
    dist.broadcast(td, 0)
    dist.broadcast(tx, 0)
    total = np.zeros((H, W), np.float32)
    valid = 0
    for i in light_shard(len(ds["lights"]), rank, world):
        err, _ = oracle.process_light(ds["lights"][i], None, td.numpy(), None, 20, None, tx.numpy(), total)
        valid += 0 if err else 1
    tsum, tval = torch.from_numpy(total), torch.tensor([valid])
    dist.reduce(tsum, 0)
    dist.reduce(tval, 0)
    if rank == 0:
        n = int(tval[0]) + 1                       # the reference frame counts once, on the root only
        out["image"] = oracle.scale(ref_cal + tsum.numpy(), n)
        out["valid"] = n
    dist.destroy_process_group()


def test_host_fed_balance_shares():
    """bench.py's split of the host-fed lights over ranks with unequal host links (balance_host_fed): every light is handed out
    exactly once, every rank keeps at least one, ranks finish together within one frame's copy time, and equal links with
    equal extras give equal shares."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    rates = [23.4] * 4 + [35.9] * 4                     # what this pool's 8-GPU box gives when all ranks copy at once
    extras = [30 / 8 + 1] + [30 / 8] * 7                # 1/8 of the calibration frames each, the reference on rank 0
    n = bench.balance_host_fed(rates, extras, 400)
    assert sum(n) == 400 and min(n) >= 1
    t = [(k + e) / r for k, e, r in zip(n, extras, rates)]
    assert max(t) - min(t) <= 1.0 / min(rates) + 1e-9   # finishing times within one frame on the slowest link
    assert bench.balance_host_fed([50.0] * 4, [2.0] * 4, 200) == [50] * 4
    assert bench.balance_host_fed([1.0, 1000.0], [0, 0], 10) == [1, 9]
    for world in (2, 3, 5, 8):
        for total in (world, 7 * world + 3, 400):
            r = [10.0 + 3 * i for i in range(world)]
            n = bench.balance_host_fed(r, [1.0] * world, total)
            assert sum(n) == total and min(n) >= 1, (world, total, n)


def test_two_gloo_ranks_reproduce_the_single_process_stack():
    ds = dataset()
    want = oracle.stack(ds["ref"], ds["lights"], 20, dark_frames=ds["darks"])
    assert want["total_valid"] >= 5
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_gloo_worker, args=(2, free_port(), out), nprocs=2, join=True)
        assert out["valid"] == want["total_valid"]
        # only the association of the fp32 sum differs (ref + (p0 + p1) vs ref + (a1 + a2 + ...))
        assert np.max(np.abs(out["image"] - want["image"])) <= 1e-6


def _nccl_worker(rank, world, port, out, shard_masters=True, full_cal=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import openskystacker_b200 as ob
    from openskystacker_b200 import binding as B
    from openskystacker_b200.distributed import init_comm, sharded_stack
    ds = dataset(full_cal)
    eng = ob.Engine(rank)
    eng.set_geometry(W, H, 1, np.uint16, 2)
    init_comm(eng, dist)
    calib = {B.DARK: ds["darks"]}
    if full_cal:
        calib.update({B.BIAS: ds["bias"], B.FLAT: ds["flats"]})
    img, tv = sharded_stack(eng, dist, ds["ref"], ds["lights"], 20, calib, shard_masters=shard_masters)
    dark = eng.get_master(B.FLAT if full_cal else B.DARK)
    darks = [None] * world
    dist.all_gather_object(darks, dark.tobytes())
    if rank == 0:
        out["darks_identical"] = all(d == darks[0] for d in darks)
        out["dark"] = dark
    reps = [(r.nstars, r.k, r.err) for r in eng.reports()]
    gathered = [None] * world
    dist.all_gather_object(gathered, reps)
    if rank == 0:
        out["image"], out["valid"], out["reports"] = img, tv, gathered
    eng.close()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world,shard_masters,full_cal", [(2, True, False), (2, False, False), (2, True, True), (4, True, True)])
def test_nccl_ranks_match_single_gpu_and_oracle(world, shard_masters, full_cal):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (run with gpurun --gpus {world})")
    ds = dataset(full_cal)
    want = oracle.stack(ds["ref"], ds["lights"], 20, ds["bias"] or None, ds["darks"], None, ds["flats"] or None)
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_nccl_worker, args=(world, free_port(), out, shard_masters, full_cal), nprocs=world, join=True)
        assert out["valid"] == want["total_valid"]
        # masters built in row slices + all-gather (or built on the root + broadcast): the same bits on every rank and in the
        # oracle — for the flat that includes the mean of channel 0 over the WHOLE frame, taken after the exchange
        o_master = want["masters"]["flat" if full_cal else "dark"]
        assert out["darks_identical"] and np.array_equal(out["dark"].view(np.uint32), o_master.view(np.uint32))
        assert np.max(np.abs(out["image"] - want["image"])) <= 1e-6
        per_rank = out["reports"]
        for r in range(world):
            idx = light_shard(len(ds["lights"]), r, world)
            assert per_rank[r] == [(want["reports"][i].nstars, want["reports"][i].k, want["reports"][i].err) for i in idx]
