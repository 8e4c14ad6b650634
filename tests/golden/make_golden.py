"""Generates the committed golden fixtures from the CPU oracle (run here, in the build container).

    python tests/golden/make_golden.py

Inputs are the deterministic synthetic textures of texture-synthesis_b200/synth.py (regenerated from
their seeds by the tests), so only the oracle's OUTPUTS are stored: final maps, resolution order and
scores, plus one frozen mid-run snapshot with per-item evaluations (k-NN list, argmin, score).
The oracle is a restatement ("parity unpinned" against the Rust binary, see oracle/ts_oracle.cpp).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.helpers import small_cases  # noqa: E402

GOLDEN_CASES = ("single_64", "multi_randinit", "inpaint_tiling", "masks_ignore", "guided")


def snapshot_items(case, n_eval=128):
    rng = np.random.RandomState(1234)
    npx = case.out_w * case.out_h
    pixels = rng.randint(0, npx, size=n_eval).astype(np.uint32)
    seeds = (np.arange(n_eval, dtype=np.uint64) + np.uint64(500)).astype(np.uint64)
    return pixels, seeds


def main():
    for case in small_cases():
        if case.name not in GOLDEN_CASES:
            continue
        g = case.run_oracle(trace=True)
        flat, score = g.resolved()
        tr = g.trace()
        out = dict(color=g.color(), coord=g.coord().astype(np.uint16), ids=g.ids(), resolved_flat=flat, resolved_score=score,
                   trace_best=tr["best"].astype(np.int16), trace_ncand=tr["ncand"].astype(np.int16))
        # frozen snapshot after ~20% of the first stages' work
        max_items = int(0.3 * case.out_w * case.out_h)
        gs = case.run_oracle(max_items=max_items)
        sflat, sscore = gs.resolved()
        pixels, seeds = snapshot_items(case)
        ev = gs.eval_items(case.oracle_params(), 1, 0.0, 4242, pixels, seeds)
        out.update(snap_max_items=np.int64(max_items), snap_color=gs.color(), snap_coord=gs.coord().astype(np.uint16), snap_ids=gs.ids(),
                   snap_tree=gs.tree_points().astype(np.int16), snap_flat=sflat, snap_score=sscore, snap_locked=np.int64(gs.locked_count()),
                   snap_pixels=pixels, snap_seeds=seeds, snap_neigh=ev["neigh"].astype(np.int32), snap_res=ev["res"], snap_escore=ev["score"])
        path = os.path.join(HERE, f"{case.name}.npz")
        np.savez_compressed(path, **out)
        print(case.name, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
