"""GPU tests (the parity tests proper): everything goes through the C ABI of libtsb200.so and is compared
bit-for-bit with the CPU oracle on the same seeded inputs, and with the committed golden fixtures."""
import os

import numpy as np
import pytest

from tests.helpers import O, Case, small_cases, compare_runs
from texture_synthesis_b200.synth import synth_texture

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def capi():
    from texture_synthesis_b200 import capi as c
    assert c.device_count() > 0, "no CUDA device: the CUDA path has no CPU fallback"
    return c


@pytest.mark.parametrize("shape", [(64, 64, 5), (100, 72, 5), (33, 47, 2), (300, 300, 5), (512, 512, 5), (17, 16, 1), (16, 16, 0)])
def test_pyramid_bit_exact(shape):
    w, h, lv = shape
    img = synth_texture(w, h, 5)
    assert (O.pyramid_build(img, lv) == capi().pyramid_build(img, lv)).all()


@pytest.mark.parametrize("filt", [0, 1, 2])
def test_resize_bit_exact(filt):
    img = synth_texture(90, 70, 9)
    for (nw, nh) in ((45, 35), (128, 96), (90, 70), (17, 200), (1, 1)):
        assert (O.resize(img, nw, nh, filt) == capi().resize(img, nw, nh, filt)).all()


def _eval_both(case, max_items, n_eval, level, alpha):
    c = capi()
    go = case.run_oracle(max_items=max_items)
    gg = case.gpu_generator()
    gg.upload_inputs(case.pyramids, case.method_list, case.mask_list, case.guides)
    fl, sc = go.resolved()
    gg.load_state(go.color(), go.coord(), go.ids(), go.tree_points(), fl, sc, go.locked_count())
    rng = np.random.RandomState(max_items % 1000 + 1)
    pixels = rng.randint(0, case.out_w * case.out_h, size=n_eval).astype(np.uint32)
    if len(fl):
        pixels[: n_eval // 2] = fl[rng.randint(0, len(fl), size=n_eval // 2)]   # redo items see themselves (q4)
    seeds = (np.arange(n_eval, dtype=np.uint64) * np.uint64(3) + np.uint64(1000)).astype(np.uint64)
    ro = go.eval_items(case.oracle_params(), level, alpha, 12345, pixels, seeds)
    rg = gg.eval_items(case.gpu_params(), level, alpha, 12345, pixels, seeds)
    return ro, rg


@pytest.mark.parametrize("case", small_cases(), ids=lambda c: c.name)
@pytest.mark.parametrize("frac", [0.0005, 0.004, 0.02, 0.2, 1.2])
def test_frozen_snapshot_argmin_bit_exact(case, frac):
    """K2+K3+K4 on frozen synthesis snapshots: same ordered k-NN list, same argmin candidate and source
    coordinate; float cost within 1e-5 relative (north_star tolerance; observed: bit-identical)."""
    case.build()
    level = 0 if frac < 0.05 else min(2, max(1, case.stages) - 1)
    alpha = 0.3 if case.guided else 0.0
    ro, rg = _eval_both(case, max(1, int(frac * case.out_w * case.out_h)), 384, level, alpha)
    assert (ro["neigh"] == rg["neigh"]).all()
    assert (ro["res"] == rg["res"]).all()
    rel = np.abs(ro["score"] - rg["score"]) / np.maximum(np.abs(ro["score"]), 1e-30)
    assert rel.max() <= 1e-5


@pytest.mark.parametrize("case", small_cases(), ids=lambda c: c.name)
def test_end_to_end_identical_to_single_thread_oracle(case):
    go = case.run_oracle()
    gg = case.run_gpu()
    r = compare_runs(go, gg)
    # mismatched-pixel fraction against the max_thread_count(1) run with the same seed: 0
    assert r["color_mismatch"] == 0 and r["coord_mismatch"] == 0 and r["id_mismatch"] == 0 and r["order_equal"], r
    assert r["score_max_rel"] <= 1e-5, r
    assert (go.uncertainty_map() == gg.uncertainty_map()).all()
    po, mo = go.id_maps()
    pg, mg = gg.id_maps()
    assert (po == pg).all() and (mo == mg).all()


@pytest.mark.parametrize("name", ["single_64", "multi_randinit", "inpaint_tiling", "masks_ignore", "guided"])
def test_against_committed_golden(name):
    case = next(c for c in small_cases() if c.name == name).build()
    gold = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    gg = case.run_gpu(trace=True)
    flat, score = gg.resolved()
    assert (gg.color() == gold["color"]).all() and (gg.coord() == gold["coord"]).all() and (gg.ids() == gold["ids"]).all()
    assert (flat == gold["resolved_flat"]).all()
    assert np.allclose(score, gold["resolved_score"], rtol=1e-5, atol=0)
    assert (gg.trace()["best"] == gold["trace_best"]).all() and (gg.trace()["ncand"] == gold["trace_ncand"]).all()
    # frozen snapshot from the fixture
    g2 = case.gpu_generator()
    g2.upload_inputs(case.pyramids, case.method_list, case.mask_list, case.guides)
    g2.load_state(gold["snap_color"], gold["snap_coord"].astype(np.uint32), gold["snap_ids"], gold["snap_tree"].astype(np.int32),
                  gold["snap_flat"], gold["snap_score"], int(gold["snap_locked"]))
    ev = g2.eval_items(case.gpu_params(), 1, 0.0, 4242, gold["snap_pixels"], gold["snap_seeds"])
    assert (ev["neigh"] == gold["snap_neigh"]).all() and (ev["res"] == gold["snap_res"]).all()
    assert np.allclose(ev["score"], gold["snap_escore"], rtol=1e-5, atol=0)


def test_medium_size_identical_and_deterministic():
    case = Case("medium", 256, 256, [(128, 128)], seed=0)
    go = case.run_oracle()
    g1 = case.run_gpu()
    r = compare_runs(go, g1)
    assert r["color_mismatch"] == 0 and r["coord_mismatch"] == 0 and r["order_equal"], r
    g2 = case.run_gpu()
    assert (g1.coord() == g2.coord()).all()


def test_reset_and_resident_rerun_reproduce():
    case = Case("rerun", 96, 96, [(64, 64)], seed=11).build()
    g = case.gpu_generator()
    g.upload_inputs(case.pyramids)
    g.resolve_resident(case.gpu_params())
    a = g.coord().copy()
    g.reset()
    g.resolve_resident(case.gpu_params())
    assert (g.coord() == a).all()


def test_full_size_properties_2048_from_512():
    """BASELINE.json headline size: size-independent invariants (the oracle would need minutes here)."""
    case = Case("headline", 2048, 2048, [(512, 512)], seed=0).build()
    g = case.gpu_generator()
    g.upload_inputs(case.pyramids)
    g.resolve_resident(case.gpu_params())
    st = g.stats()
    assert st["work_items"] == 8257536                                   # 1.96875 * 2048^2
    flat, score = g.resolved()
    assert len(flat) == 2048 * 2048 and len(np.unique(flat)) == 2048 * 2048
    co, col = g.coord(), g.color()
    assert co[..., 0].max() < 512 and co[..., 1].max() < 512 and co[..., 2].max() == 0
    assert (col == case.pyramids[0][-1][co[..., 1], co[..., 0]]).all()   # colour == example[coord]
    assert np.isfinite(score).all() and (score >= 0).all()
    assert st["texels_fetched"] <= st["texels_nominal"] == st["candidates"] * 50 - (st["candidates"] * 50 - st["texels_nominal"])
    # coherence: most neighbouring output pixels continue the same source patch
    dx = (co[:, 1:, 0].astype(int) - co[:, :-1, 0].astype(int) == 1) & (co[:, 1:, 1] == co[:, :-1, 1])
    assert dx.mean() > 0.5


def test_session_mirror_runs_like_reference_example_01():
    import texture_synthesis_b200 as ts
    ex = synth_texture(64, 64, 1)
    seen = []
    gen = ts.Session.builder().add_example(ex).seed(120).output_size(ts.Dims.square(100)).max_thread_count(1).build() \
        .run(lambda img, total, stage: seen.append(total))
    img = gen.into_image()
    assert img.shape == (100, 100, 4) and (img[..., 3] == 255).all()
    assert seen and seen[-1][0] <= seen[-1][1]
    ct = gen.get_coordinate_transform()
    assert (ct.apply([ex]) == img).all()                                 # repeat_transform (lib/tests/diff.rs:254-284)
    # the same session expressed directly on the oracle
    case_pyr = O.pyramid_build(ex, 5)
    go = O.Generator(100, 100)
    go.set_examples([case_pyr])
    go.resolve(O.make_params(seed=120))
    assert (go.color() == img).all()


@pytest.mark.parametrize("env", [{"TSB_CHUNK": "700"}, {"TSB_CHUNK": "2000", "TSB_RING_ITEMS": "4500"}, {"TSB_RING_ITEMS": "900"},
                                 {"TSB_NO_FAST": "1"}, {"TSB_BRUTE_BELOW": "0"}, {"TSB_NO_L2_WINDOW": "1"}],
                         ids=["small_chunks", "ring_wraps", "tiny_ring_resplit", "general_scoring", "mask_search_only", "no_l2_window"])
def test_scheduler_variants_give_the_same_result(env, monkeypatch):
    """The streaming scheduler cuts every phase into chunks whose lists live in a ring that the analysis stream refills
    as the resolve stream frees it; chunk size, ring size (wrap-around, re-splitting), the scoring path and the k-NN search
    variant must not change a single bit of the result."""
    cases = [Case("sched", 96, 80, [(64, 48)], seed=17, tiling=True).build(),
             Case("sched_big_tiling", 208, 176, [(64, 48)], seed=18, tiling=True).build(),
             Case("sched_big", 200, 168, [(72, 56)], seed=19).build()]
    refs = [c.run_gpu() for c in cases]
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    for case, ref in zip(cases, refs):
        alt = case.run_gpu()
        assert (ref.coord() == alt.coord()).all() and (ref.color() == alt.color()).all(), case.name
        fa, sa = ref.resolved()
        fb, sb = alt.resolved()
        assert (fa == fb).all() and (sa.view(np.uint32) == sb.view(np.uint32)).all(), case.name


@pytest.mark.parametrize("shape", [(48, 40), (128, 96), (400, 382)])
def test_guide_preprocessing_bit_exact(shape):
    """N2: blur(sigma=2) -> grayscale and histogram matching (utils.rs:101-183) against the oracle."""
    w, h = shape
    a, b = synth_texture(w, h, 3), synth_texture(64, 80, 7)
    g_o, g_g = O.guide_map(a, 2.0), capi().guide_map(a, 2.0)
    assert (g_o == g_g).all()
    t_o = O.guide_map(b, 2.0)
    assert (O.match_histograms(g_o, t_o) == capi().match_histograms(g_g, t_o)).all()


def test_style_transfer_session_matches_oracle():
    """Config 3 (style transfer auto-guides, lib/examples/04_style_transfer.rs) through the Session mirror."""
    import texture_synthesis_b200 as ts
    ex, tgt = synth_texture(72, 72, 11), synth_texture(90, 90, 12)
    gen = ts.Session.builder().add_example(ex).load_target_guide(tgt).output_size(ts.Dims.square(90)).seed(5).build().run(None)
    img = gen.into_image()
    # the same pipeline on the oracle
    tg = O.guide_map(tgt, 2.0)
    pyr, tpyr = O.pyramid_build(ex, 5), O.pyramid_build(tg, 5)
    eg = O.pyramid_build(O.match_histograms(O.guide_map(ex, 2.0), tg), 5)
    go = O.Generator(90, 90)
    go.set_examples([pyr])
    go.set_guides(tpyr, [eg])
    go.resolve(O.make_params(seed=5))
    assert (go.color() == img).all()


from tests.helpers import edge_cases  # noqa: E402


@pytest.mark.parametrize("case", edge_cases(), ids=lambda c: c.name)
def test_edge_cases_identical_to_oracle(case):
    go, gg = case.run_oracle(), case.run_gpu()
    r = compare_runs(go, gg, check_scores=False)
    assert r["color_mismatch"] == 0 and r["coord_mismatch"] == 0 and r["id_mismatch"] == 0 and r["order_equal"], r
    so, sg = go.resolved()[1], gg.resolved()[1]
    assert (np.isnan(so) == np.isnan(sg)).all()
    ok = ~np.isnan(so) & ~np.isinf(so)
    assert (np.isinf(so) == np.isinf(sg)).all()
    assert np.allclose(so[ok], sg[ok], rtol=1e-5, atol=0)


def test_unsupported_and_invalid_inputs_fail_loudly():
    c = capi()
    case = Case("bad", 32, 32, [(16, 16)], seed=0).build()
    g = case.gpu_generator()
    for kw in (dict(k=129), dict(k=100, m=200), dict(cauchy=1.5), dict(m=0), dict(p=-0.1)):
        with pytest.raises(c.TsbError):
            g.resolve(c.make_params(**kw), case.pyramids)
    with pytest.raises(c.TsbError):       # pyramid too shallow for the requested stages
        g.resolve(c.make_params(stages=5), [case.pyramids[0][:2]])
    with pytest.raises(c.TsbError):       # image too small for 5 levels (the reference would panic on an empty image)
        c.pyramid_build(synth_texture(8, 8, 1), 5)
    all_zero = np.zeros((16, 16, 4), np.uint8)
    with pytest.raises(c.TsbError):       # a sampling mask that allows nothing would loop forever in the reference
        g.resolve(c.make_params(), case.pyramids, [c.SAMPLE_IMAGE], [all_zero])


def test_progress_callback_polls_without_changing_the_result():
    """ProgressNotifier (ms.rs:1054-1107): the calling thread polls the device's claim counter, as the reference's main thread
    polls its atomic (ms.rs:1026-1034), and reports each change of the integer percentage with a snapshot of the colours; no
    kernel is stalled or split for it, and the result is bit-identical to a run without a callback."""
    import threading
    case = Case("progress", 1536, 1536, [(256, 256)], seed=4).build()   # long enough (~30 ms) for the poll to see many percentages
    ref = case.run_gpu()
    calls = []
    me = threading.get_ident()

    def cb(img, total, stage):
        assert threading.get_ident() == me           # on the caller's thread only (Box<dyn GeneratorProgress> is not Send)
        assert img.shape == (1536, 1536, 4)
        calls.append((total[0], total[1], stage[0], stage[1], int(img[..., 3].max())))
    g = case.gpu_generator()
    g.resolve(case.gpu_params(), case.pyramids, case.method_list, case.mask_list, case.guides, progress=cb)
    assert (g.coord() == ref.coord()).all() and (g.color() == ref.color()).all()
    # With a cheap callback 75-85 calls arrive at every size from 256^2 to 2048^2 (tests/gpu_progress_count.py; the reference:
    # <= 101).  This one scans 9 MB per call, so -- as with the reference's spinning main thread -- percentages pass meanwhile.
    assert len(calls) >= 10
    cur = [c[0] for c in calls]
    assert cur == sorted(cur) and round(100.0 * calls[-1][0] / calls[-1][1]) == 100   # monotonic, ends at 100 %
    assert all(c[1] == calls[0][1] for c in calls)
    assert all(0 <= c[2] <= c[3] for c in calls)
    pcnt = [round(100.0 * c[0] / c[1]) for c in calls]
    assert len(set(pcnt)) == len(pcnt)                                      # one call per integer percentage at most
    assert calls[-1][4] == 255                                              # the snapshot carries real colours


def test_watchdog_reports_a_stall_instead_of_hanging(monkeypatch):
    """A work item that waits longer than TSB_WATCHDOG_MS (wall clock, %globaltimer) aborts the run with an error; the
    limit is generous by default (20 s) so that sanitizers, debuggers and time slicing do not trip it."""
    case = Case("wd", 64, 64, [(32, 32)], seed=1).build()
    monkeypatch.setenv("TSB_WATCHDOG_MS", "20000")
    g = case.run_gpu()
    assert g.stats()["work_items"] > 0
