"""Small end-to-end runs for compute-sanitizer (not a pytest file):
    compute-sanitizer --tool memcheck  python tests/gpu_sanitize.py
    compute-sanitizer --tool racecheck python tests/gpu_sanitize.py
Every configuration runs through the C ABI and is compared with the CPU oracle, so a clean sanitizer log belongs to runs
that completed with the right result.  TSB_WATCHDOG_MS is raised: under a sanitizer a kernel runs ~100x slower."""
import os
import sys

os.environ.setdefault("TSB_WATCHDOG_MS", "600000")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import Case, compare_runs, O  # noqa: E402

CASES = [
    Case("plain", 48, 48, [(32, 32)], seed=3),
    Case("tiling", 48, 48, [(32, 32)], seed=7, tiling=True),
    Case("inpaint_tiling", 48, 48, [(48, 48)], seed=5, tiling=True, inpaint=True),
    Case("guided", 40, 40, [(32, 32)], seed=2, guided=True),
    Case("multi_masks", 40, 40, [(24, 24), (20, 20), (28, 24)], seed=211,
         methods=[O.METHOD_IMAGE, O.METHOD_IGNORE, O.METHOD_ALL], sample_masks=True),
    Case("multi_randinit", 40, 40, [(24, 24), (20, 20), (28, 24)], seed=211, random_init=6),
    Case("cauchy0", 24, 24, [(16, 16)], seed=12, cauchy=0.0, stages=2, k=10, m=6),
]
if len(sys.argv) > 1:
    CASES = [c for c in CASES if c.name in sys.argv[1:]]
bad = 0
for case in CASES:
    go, gg = case.run_oracle(), case.run_gpu()
    r = compare_runs(go, gg, check_scores=False)
    ok = r["color_mismatch"] == 0 and r["coord_mismatch"] == 0 and r["id_mismatch"] == 0 and r["order_equal"]
    print(f"[sanitize] {case.name}: {'identical to the oracle' if ok else 'MISMATCH ' + str(r)}", flush=True)
    bad += 0 if ok else 1
sys.exit(1 if bad else 0)
