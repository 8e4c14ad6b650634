"""Band-sharded multi-GPU check (run under torchrun, one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/gpu_mg_check.py [OUT EX]
Every rank runs the same synthesis linked through CUDA IPC; the result must be bit-identical to a single-GPU run."""
import hashlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from tests.helpers import Case
from texture_synthesis_b200 import capi
from texture_synthesis_b200.parallel import link_band_sharded

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
out = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ex = int(sys.argv[2]) if len(sys.argv) > 2 else 128
tiling = len(sys.argv) > 3 and sys.argv[3] == "tiling"
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
case = Case(f"mg_{out}", out, out, [(ex, ex)], seed=0, tiling=tiling).build()
params = case.gpu_params()

g = capi.Generator(out, out, device=local)
g.upload_inputs(case.pyramids)
link_band_sharded(g, params, dist)
times = []
for it in range(2):
    g.reset()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    g.resolve_resident(params)
    torch.cuda.synchronize(); dist.barrier()
    times.append(time.perf_counter() - t0)
coord = g.coord()
color = g.color()
flat, score = g.resolved()
digest = hashlib.sha256(coord.tobytes() + color.tobytes() + score.tobytes()).hexdigest()
all_digests = [None] * world
dist.all_gather_object(all_digests, digest)
st = g.stats()
if rank == 0:
    print(f"[mg] world {world} out {out}^2: {times[-1]*1e3:.1f} ms ({out*out/times[-1]/1e6:.2f} Mpx/s), band-sharded phases {g.mg_phases()}, "
          f"replicas identical: {len(set(all_digests)) == 1}", flush=True)
    # single-GPU reference on the same device
    g1 = capi.Generator(out, out, device=local)
    g1.upload_inputs(case.pyramids)
    g1.resolve_resident(params)
    t0 = time.perf_counter(); g1.reset(); g1.resolve_resident(params); t1 = time.perf_counter() - t0
    same = bool((g1.coord() == coord).all() and (g1.color() == color).all())
    f1, s1 = g1.resolved()
    same = same and bool((f1 == flat).all() and (s1.view(np.uint32) == score.view(np.uint32)).all())
    print(f"[mg] single GPU: {t1*1e3:.1f} ms; multi-GPU result identical to single-GPU: {same}", flush=True)
    print(f"[mg] rank0 stats: resolve_ms {st['gpu_ms_resolve']:.1f} analysis_ms {st['gpu_ms_analysis']:.1f} phases {st['phases']}", flush=True)
dist.barrier()
dist.destroy_process_group()
