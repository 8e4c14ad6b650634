"""Randomised parity sweep (not a pytest file): python tests/gpu_fuzz.py [n_cases] [seed]

Draws synthesis configurations at random (sizes, k, m, stages, backtrack, tiling, inpaint, several examples, sampling
masks, guides, random_init, examples with non-opaque alpha) and compares the CUDA path, through the C ABI, with the
single-thread run of the CPU oracle: colour / coordinate / id maps, resolution order and scores must be identical.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ts_oracle as O  # noqa: E402
from tests.helpers import Case, compare_runs  # noqa: E402


def random_case(rng, i):
    out_w, out_h = int(rng.integers(40, 260)), int(rng.integers(40, 260))
    kind = rng.choice(["plain", "tiling", "inpaint", "inpaint_tiling", "multi", "masks", "guided", "randinit"])
    k = int(rng.choice([1, 5, 20, 50, 50, 50, 80]))
    m = int(rng.choice([1, 10, 50, 50, 100]))  # the reference rejects m = 0 (session.rs:489)
    stages = int(rng.choice([0, 1, 3, 5, 5, 5]))
    p = float(rng.choice([0.3, 0.5, 0.5, 0.8]))
    kw = dict(seed=int(rng.integers(0, 1 << 30)), k=k, m=m, stages=stages, p=p,
              cauchy=float(rng.choice([0.25, 0.5, 1.0])), tex_seed=int(rng.integers(1, 1000)))
    if k + kw["m"] > 250:
        kw["m"] = 250 - k
    def ex():
        return (int(rng.integers(24, 120)), int(rng.integers(24, 120)))
    if kind == "plain":
        return Case(f"fz{i}_plain", out_w, out_h, [ex()], **kw)
    if kind == "tiling":
        return Case(f"fz{i}_tiling", out_w, out_h, [ex()], tiling=True, **kw)
    if kind in ("inpaint", "inpaint_tiling"):
        w, h = min(out_w, 160), min(out_h, 160)
        return Case(f"fz{i}_{kind}", w, h, [(w, h)], inpaint=True, tiling=kind.endswith("tiling"), **kw)
    if kind == "multi":
        return Case(f"fz{i}_multi", out_w, out_h, [ex(), ex(), ex()], **kw)
    if kind == "masks":
        return Case(f"fz{i}_masks", out_w, out_h, [ex(), ex(), ex()], sample_masks=True,
                    methods=[O.METHOD_IMAGE, O.METHOD_IGNORE, O.METHOD_ALL], **kw)
    if kind == "guided":
        return Case(f"fz{i}_guided", out_w, out_h, [ex()], guided=True, alpha=float(rng.choice([0.3, 0.8])), **kw)
    return Case(f"fz{i}_randinit", out_w, out_h, [ex(), ex()], random_init=int(rng.integers(1, 200)), **kw)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    bad = 0
    t0 = time.time()
    for i in range(n):
        case = random_case(rng, i).build()
        if rng.random() < 0.25:  # non-opaque inputs exercise the alpha channel as a full cost channel (q8)
            for e in case.examples:
                e[..., 3] = rng.integers(0, 256, e.shape[:2], dtype=np.uint8)
            case.pyramids = [O.pyramid_build(e, max(1, case.stages)) for e in case.examples]
            if case.inpaint:
                case.inpaint_color = case.examples[0].copy()
        try:
            go = case.run_oracle()
            gg = case.run_gpu()
            st = compare_runs(go, gg)
            ok = (st["color_mismatch"] == 0 and st["coord_mismatch"] == 0 and st["id_mismatch"] == 0 and st["order_equal"]
                  and st.get("score_bit_mismatch", 0) == 0)
        except Exception as exc:  # a refusal (unsupported parameters) is reported, not counted as a mismatch
            print(f"[fuzz] {case.name} {case.out_w}x{case.out_h} k={case.k} m={case.m} stages={case.stages}: {type(exc).__name__}: {exc}")
            continue
        if not ok:
            bad += 1
        print(f"[fuzz] {case.name} {case.out_w}x{case.out_h} ex={case.ex_sizes} k={case.k} m={case.m} stages={case.stages} p={case.p}: "
              f"{'identical' if ok else 'MISMATCH ' + str(st)}", flush=True)
    print(f"[fuzz] {n} cases, {bad} mismatching, {time.time() - t0:.1f} s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
