"""Multi-GPU sharding of a batch of tracks (one process per GPU).

Tracks are independent units keyed by their id: rank r of W owns a contiguous
range of ids and draws the Philox streams (seed, id), so a W-GPU run transports
exactly the tracks a single-GPU run of the same total size would. There is no
data-path collective; the only exchange is one all-reduce (sum) of the tally
vector per batch (NCCL on GPUs, gloo in the CPU tests)."""


def shard_range(n_total, rank, world):
    """Half-open range of track indices owned by `rank`."""
    base, extra = divmod(n_total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def step_offset(step, rank, world, n_per_rank):
    """Id of the first track of `rank` at step `step` (weak scaling: every
    rank transports n_per_rank fresh tracks per step)."""
    return (step * world + rank) * n_per_rank


def allreduce_tallies(tensor):
    """Sum the tally vector over the ranks, in place. No-op without a group."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and \
            dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor
