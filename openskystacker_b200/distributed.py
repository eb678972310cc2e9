"""Multi-GPU plumbing: one process per GPU, frames sharded across ranks exactly like the
reference shards them across `-j` worker threads (libstacker/src/util.cpp:738: worker i
takes lights i, i+N, i+2N, ...), masters + reference broadcast from rank 0, and ONE
ncclReduce of the float32 sum image + the valid-frame count back to rank 0
(libstacker/src/imagestacker.cpp:348-351 is the reference's serial version of that reduce).
torch.distributed is used only to ship the 128-byte NCCL id and for barriers."""
from . import binding as B


def light_shard(nlights, rank, world):
    """Indices of the lights rank `rank` of `world` processes: the reference's strided split."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return list(range(rank, nlights, world))


def init_comm(eng, dist):
    """Create the engine's NCCL communicator over the ranks of torch.distributed's default group."""
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [B.Engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    eng.comm_init(box[0], rank, world)
    return rank, world


def sharded_stack(eng, dist, ref, lights, threshold, calib=None, mem=B.HOST, want_image=True, shard_masters=True):
    """Whole stack over all ranks.  `lights` is the FULL list on every rank (each rank picks its shard); `ref` is only
    read on rank 0.  `calib` (dict kind -> frames): with shard_masters every rank has it and accumulates ITS row slice of
    every calibration frame, the slices are exchanged over NVLink (oss_comm_allgather_master; bit-identical masters);
    otherwise only rank 0 reads it, builds the masters and broadcasts them.  Returns (image, total_valid) on rank 0 and
    (None, 0) elsewhere."""
    rank, world = dist.get_rank(), dist.get_world_size()
    have = rank == 0 or shard_masters
    kinds = [[k for k in (B.BIAS, B.DARK, B.DARKFLAT, B.FLAT) if calib and calib.get(k)] if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(kinds, src=0)     # which masters exist (every rank of a real job reads it off the image list)
    eng.clear_masters()
    for kind in kinds[0]:                            # imagestacker.cpp:262-281 order
        if world > 1 and shard_masters:
            row0, rows = eng.comm_master_slice()
            eng.build_master_slice(kind, calib[kind], row0, rows, mem)
            eng.comm_allgather_master(kind)          # asynchronous: overlaps the next master's slice and the reference
            continue
        if rank == 0 and have:
            eng.build_master(kind, calib[kind], mem)
        if world > 1:
            eng.comm_broadcast_master(kind, 0)       # asynchronous: overlaps the next master and the reference on rank 0
    if rank == 0:
        eng.set_reference(ref, threshold, mem)
    if world > 1:
        eng.comm_broadcast_reference(threshold, 0)   # nobody waits on the host
    mine = [lights[i] for i in light_shard(len(lights), rank, world)]
    keep = None
    if mine:
        ptrs, keep = eng._ptrs(mine, mem)
        eng.add_light_ptrs(ptrs, len(mine), mem)     # asynchronous; `keep` holds the host frames until the stack is drained
    if world > 1:
        eng.comm_reduce(0)                           # the last batch's warp is still pending: banded warp + reduce overlap
    if rank == 0:
        out = eng.finish(want_image)
        del keep
        return out
    eng.sync()
    del keep
    return None, 0
