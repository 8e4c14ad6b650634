"""Timing driver (not a pytest file): python tests/gpu_perf.py OUT EX [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import Case
out, ex = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
case = Case(f"perf_{out}", out, out, [(ex, ex)], seed=0).build()
g = case.gpu_generator()
g.upload_inputs(case.pyramids)
for it in range(reps):
    g.reset()
    t0 = time.time()
    g.resolve_resident(case.gpu_params())
    dt = time.time() - t0
    st = g.stats()
    print(f"perf {out}^2 from {ex}^2: {dt * 1e3:.1f} ms -> {out * out / dt / 1e6:.3f} Mpx/s; rounds {st['rounds']} phases {st['phases']} "
          f"resolve_ms {st['gpu_ms_resolve']:.1f} analysis_ms {st['gpu_ms_analysis']:.1f} sched_ms {st['host_ms_schedule']:.1f} "
          f"texels {st['texels_fetched'] / 1e9:.2f}G of nominal {st['texels_nominal'] / 1e9:.2f}G", flush=True)
