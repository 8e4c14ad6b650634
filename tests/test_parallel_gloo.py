"""CPU test of the N>1 path's host logic with world_size 2 over gloo (the GPU run uses the same code with NCCL)."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from texture_synthesis_b200.parallel import shard_sessions, aggregate_throughput, max_over_ranks
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_sessions(5, rank, world)
    # each rank "processes" its sessions: rank 0 is slower
    secs = 2.0 if rank == 0 else 1.0
    thr, total, t = aggregate_throughput(len(mine) * 100.0, secs, dist)
    dist.barrier()
    q.put((rank, mine, thr, total, t, max_over_ranks(rank, dist)))
    dist.destroy_process_group()


def test_session_sharding_and_max_over_ranks_world2():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, thr0, tot0, t0, mx0), (r1, m1, thr1, tot1, t1, mx1) = res
    assert m0 == [0, 1, 2] and m1 == [3, 4]                     # balanced, contiguous, complete
    assert tot0 == tot1 == 500.0 and t0 == t1 == 2.0             # sum of units, MAX of the per-rank times
    assert thr0 == thr1 == 250.0 and mx0 == mx1 == 1.0


def test_shard_sessions_covers_everything():
    sys.path.insert(0, ROOT)
    from texture_synthesis_b200.parallel import shard_sessions
    for n in (0, 1, 7, 8, 9, 64):
        for world in (1, 2, 4, 8):
            got = sum((shard_sessions(n, r, world) for r in range(world)), [])
            assert got == list(range(n))
