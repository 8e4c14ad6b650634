"""Staged bring-up checks on a GPU box (not a pytest file): prints where the CUDA path and the oracle diverge.

usage: python tests/gpu_debug.py [stage ...]   stages: micro pyramid eval e2e perf
"""
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import Case, small_cases, compare_runs, O  # noqa: E402
from texture_synthesis_b200 import capi  # noqa: E402
from texture_synthesis_b200.synth import synth_texture  # noqa: E402


def stage_micro():
    for nbytes in (1 << 20, 4 << 20, 64 << 20):
        for mode in (0, 1):
            gbs, gps = capi.microbench_gather(nbytes, mode, 20)
            print(f"gather window={nbytes >> 20}MiB mode={'tex' if mode else 'ldg'}: {gps / 1e9:.1f} Ggather/s  {gbs:.1f} GB/s useful")


def stage_pyramid():
    for (w, h, lv) in ((64, 64, 5), (100, 72, 5), (48, 48, 3), (300, 300, 5), (33, 47, 2), (512, 512, 5)):
        img = synth_texture(w, h, 5)
        a = O.pyramid_build(img, lv)
        b = capi.pyramid_build(img, lv)
        print(f"pyramid {w}x{h} L{lv}: mismatching bytes {int((a != b).sum())} of {a.size}")
    img = synth_texture(90, 70, 9)
    for f in (0, 1, 2):
        for (nw, nh) in ((45, 35), (128, 96), (90, 70), (17, 200)):
            a = O.resize(img, nw, nh, f)
            b = capi.resize(img, nw, nh, f)
            print(f"resize filter {f} -> {nw}x{nh}: mismatching bytes {int((a != b).sum())}")


def eval_compare(case, max_items, n_eval=512, level=0, alpha=0.0):
    go = case.run_oracle(max_items=max_items)
    gg = case.gpu_generator()
    gg.upload_inputs(case.pyramids, case.method_list, case.mask_list, case.guides)
    fl, sc = go.resolved()
    gg.load_state(go.color(), go.coord(), go.ids(), go.tree_points(), fl, sc, go.locked_count())
    rng = np.random.RandomState(max_items % 1000 + 1)
    npx = case.out_w * case.out_h
    pixels = rng.randint(0, npx, size=n_eval).astype(np.uint32)
    if len(fl):  # half of them already-resolved pixels (redo items see themselves, quirk q4)
        pixels[: n_eval // 2] = fl[rng.randint(0, len(fl), size=n_eval // 2)]
    seeds = (np.arange(n_eval, dtype=np.uint64) * np.uint64(3) + np.uint64(1000)).astype(np.uint64)
    seeds[: n_eval // 4] = np.arange(n_eval // 4, dtype=np.uint64) + np.uint64(77)
    p_seed = 12345
    ro = go.eval_items(case.oracle_params(), level, alpha, p_seed, pixels, seeds)
    rg = gg.eval_items(case.gpu_params(), level, alpha, p_seed, pixels, seeds)
    neigh_bad = (ro["neigh"] != rg["neigh"]).any(axis=(1, 2))
    res_bad = (ro["res"] != rg["res"]).any(axis=1)
    so, sg = ro["score"], rg["score"]
    bit_bad = so.view(np.uint32) != sg.view(np.uint32)
    rel = np.abs(so - sg) / np.maximum(np.abs(so), 1e-30)
    print(f"  eval {case.name} after {max_items} items (resolved {len(fl)}): neigh mismatch {int(neigh_bad.sum())}/{n_eval}, "
          f"res mismatch {int(res_bad.sum())}, score bit mismatch {int(bit_bad.sum())}, max rel {float(rel.max()):.3e}")
    if neigh_bad.any():
        i = int(np.argmax(neigh_bad))
        print("   first neigh mismatch item", i, "pixel", pixels[i], (pixels[i] % case.out_w, pixels[i] // case.out_w))
        print("   oracle", ro["neigh"][i][:8].tolist(), "n", ro["res"][i][0])
        print("   gpu   ", rg["neigh"][i][:8].tolist(), "n", rg["res"][i][0])
    elif res_bad.any():
        i = int(np.argmax(res_bad))
        print("   first res mismatch item", i, "oracle", ro["res"][i].tolist(), so[i], "gpu", rg["res"][i].tolist(), sg[i])
    return int(neigh_bad.sum()) + int(res_bad.sum())


def stage_eval():
    for case in small_cases():
        case.build()
        total = case.out_w * case.out_h
        for frac in (0.0005, 0.004, 0.02, 0.2, 1.2):
            try:
                eval_compare(case, max(1, int(frac * total)))
            except Exception:
                traceback.print_exc()


def stage_e2e():
    for case in small_cases():
        try:
            t0 = time.time()
            go = case.run_oracle(trace=True)
            t1 = time.time()
            gg = case.run_gpu(trace=True)
            t2 = time.time()
            cmp_ = compare_runs(go, gg)
            st = gg.stats()
            print(f"e2e {case.name}: oracle {t1 - t0:.2f}s gpu {t2 - t1:.2f}s {cmp_}")
            print(f"    rounds {st['rounds']} phases {st['phases']} launches {st['kernel_launches']} resolve_ms {st['gpu_ms_resolve']:.1f} "
                  f"analysis_ms {st['gpu_ms_analysis']:.1f} sched_ms {st['host_ms_schedule']:.1f} wall_ms {st['wall_ms_total']:.1f}")
            if cmp_["color_mismatch"] or cmp_["coord_mismatch"] or not cmp_["order_equal"]:
                to, tg = go.trace(), gg.trace()
                n = min(len(to["pixel"]), len(tg["pixel"]))
                print("    trace lengths", len(to["pixel"]), len(tg["pixel"]))
                bad = (to["pixel"][:n] != tg["pixel"][:n]) | (to["best"][:n] != tg["best"][:n]) | (to["ncand"][:n] != tg["ncand"][:n]) | \
                      (to["nneigh"][:n] != tg["nneigh"][:n])
                if bad.any():
                    i = int(np.argmax(bad))
                    print(f"    first trace divergence at item {i} of {n} ({int(bad.sum())} differ):")
                    for key in ("pixel", "best", "ncand", "nneigh", "score"):
                        print("      ", key, "oracle", to[key][max(0, i - 1): i + 3].tolist(), "gpu", tg[key][max(0, i - 1): i + 3].tolist())
        except Exception:
            traceback.print_exc()


def stage_perf():
    for (out, ex) in ((256, 128), (512, 256), (1024, 512)):
        case = Case(f"perf_{out}", out, out, [(ex, ex)], seed=0).build()
        g = case.gpu_generator()
        g.upload_inputs(case.pyramids)
        for it in range(2):
            g.reset()
            t0 = time.time()
            g.resolve_resident(case.gpu_params())
            dt = time.time() - t0
        st = g.stats()
        print(f"perf {out}^2 from {ex}^2: {dt * 1e3:.1f} ms -> {out * out / dt / 1e6:.3f} Mpx/s; rounds {st['rounds']} phases {st['phases']} "
              f"resolve_ms {st['gpu_ms_resolve']:.1f} analysis_ms {st['gpu_ms_analysis']:.1f} sched_ms {st['host_ms_schedule']:.1f} "
              f"texels {st['texels_fetched'] / 1e9:.2f}G of nominal {st['texels_nominal'] / 1e9:.2f}G")


if __name__ == "__main__":
    stages = sys.argv[1:] or ["micro", "pyramid", "eval", "e2e", "perf"]
    print("devices:", capi.device_count())
    for s in stages:
        print(f"==== {s} ====", flush=True)
        t = time.time()
        try:
            globals()["stage_" + s]()
        except Exception:
            traceback.print_exc()
        print(f"==== {s} done in {time.time() - t:.1f}s ====", flush=True)
