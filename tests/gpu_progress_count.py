"""Measurement (not a pytest file): how many progress callbacks a run gets and what they cost.
python tests/gpu_progress_count.py [sizes...]   -- prints callbacks, distinct percentages and the time with / without a callback."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import Case
sizes = [int(a) for a in sys.argv[1:]] or [256, 512, 1024, 2048]
for out in sizes:
    ex = max(64, out // 4)
    case = Case(f"prog_{out}", out, out, [(ex, ex)], seed=0).build()
    g = case.gpu_generator()
    g.upload_inputs(case.pyramids)
    calls = []
    def cb(img, total, stage):
        calls.append(total[0])
    t_plain, t_cb = [], []
    for it in range(4):
        g.reset(); t0 = time.perf_counter(); g.resolve_resident(case.gpu_params()); t_plain.append(time.perf_counter() - t0)
        g.reset(); del calls[:]; t0 = time.perf_counter(); g.resolve_resident(case.gpu_params(), progress=cb); t_cb.append(time.perf_counter() - t0)
    print(f"progress {out}^2 from {ex}^2: {len(calls)} callbacks; run {min(t_plain) * 1e3:.2f} ms without, {min(t_cb) * 1e3:.2f} ms with "
          f"(+{(min(t_cb) / min(t_plain) - 1) * 100:.1f} %)", flush=True)
