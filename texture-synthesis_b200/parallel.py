"""Multi-GPU plumbing for batch mode: independent synthesis sessions sharded over ranks (one process per GPU).

The hot path shards by session ("independent sessions (batch): replicas only"): there is no
data-path collective, only a barrier and a max-over-ranks reduction of the device time.  `torch.distributed`
is plumbing (NCCL on GPUs; gloo in the CPU tests).
"""


def shard_sessions(n_sessions, rank, world):
    """Contiguous, balanced assignment of session indices to ranks (first `n % world` ranks get one more)."""
    base, extra = divmod(n_sessions, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def max_over_ranks(value, dist=None, device="cpu"):
    """Max of a per-rank scalar (device time of the timed region) over all ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, dist=None, device="cpu"):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_this_rank, seconds_this_rank, dist=None, device="cpu"):
    """Whole-job throughput: units processed by all ranks / max over ranks of the timed region."""
    total = sum_over_ranks(units_this_rank, dist, device)
    t = max_over_ranks(seconds_this_rank, dist, device)
    return total / t, total, t


def link_band_sharded(generator, params, dist):
    """Links the replicas of one output across the ranks of a node (one process per GPU) for band-sharded
    execution: CUDA IPC handles are exchanged with torch.distributed, the barrier is torch.distributed.barrier."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = generator.mg_export(params)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()
    generator.mg_attach(rank, world, gathered, barrier)
    dist.barrier()
    return rank, world
