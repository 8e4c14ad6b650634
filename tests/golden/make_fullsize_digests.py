"""Runs the CPU oracle (1 thread, the analogue of max_thread_count(1), lib/tests/diff.rs:143-145) OFFLINE on every
BASELINE.json configuration at its stated size and on the reference's nine integration configurations at the
reference's sizes, and commits SHA-256 digests of everything the reference hands back (colour, coordinate
transform, ids, resolution order, scores; map_id/patch_id debug images where the config asks for them).

    python tests/golden/make_fullsize_digests.py [case ...]        # default: every case not yet in the JSON
    python tests/golden/make_fullsize_digests.py --all             # recompute everything (C5 takes about an hour)

tests/test_gpu_fullsize.py compares the CUDA path against tests/golden/fullsize_digests.json; bench.py checks the
headline digest after its timed region.  The cases themselves are defined in tests/fullsize_cases.py.
"""
import fcntl
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests import fullsize_cases as F  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    have = F.load_digests()
    names = args or [n for n in F.SPECS if "--all" in sys.argv or n not in have]
    for name in names:
        spec = F.SPECS[name]()
        t0 = time.time()
        g = F.to_oracle(spec).run()
        d = F.digest_of_oracle(g, spec)
        d["oracle_seconds"] = round(time.time() - t0, 1)
        print(name, d, flush=True)
        with open(F.DIGESTS + ".lock", "w") as lk:  # several instances may run side by side
            fcntl.flock(lk, fcntl.LOCK_EX)
            have = F.load_digests()
            have[name] = d
            tmp = F.DIGESTS + f".tmp{os.getpid()}"
            with open(tmp, "w") as f:
                json.dump(have, f, indent=1, sort_keys=True)
            os.replace(tmp, F.DIGESTS)


if __name__ == "__main__":
    main()
