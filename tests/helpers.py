"""Shared builders for the parity tests: the same seeded inputs go to the oracle and to the CUDA path."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ts_oracle as O  # noqa: E402  (test infrastructure)
from texture_synthesis_b200.synth import synth_texture, border_inpaint_mask  # noqa: E402


def gpu_available():
    try:
        from texture_synthesis_b200 import capi
        return capi.device_count() > 0
    except Exception:
        return False


class Case:
    """One synthesis configuration, expressed in the terms of Session::builder()."""

    def __init__(self, name, out_w, out_h, ex_sizes, seed=0, k=50, m=50, stages=5, p=0.5, cauchy=1.0, alpha=0.8,
                 tiling=False, random_init=0, inpaint=False, methods=None, sample_masks=False, guided=False, tex_seed=1,
                 guide_sizes=None):
        self.name = name
        self.out_w, self.out_h = out_w, out_h
        self.ex_sizes = ex_sizes
        self.seed, self.k, self.m, self.stages, self.p, self.cauchy, self.alpha = seed, k, m, stages, p, cauchy, alpha
        self.tiling, self.random_init, self.inpaint = tiling, random_init, inpaint
        self.methods = methods
        self.sample_masks = sample_masks
        self.guided = guided
        self.guide_sizes = guide_sizes   # example-guide sizes when they differ from the examples' (the reference keeps them as loaded)
        self.tex_seed = tex_seed
        self._built = False

    def build(self):
        if self._built:
            return self
        levels = max(1, self.stages)
        self.examples = [synth_texture(w, h, self.tex_seed + i) for i, (w, h) in enumerate(self.ex_sizes)]
        if self.inpaint:
            # inpaint forces example and output to the same size (session.rs:346-379)
            assert self.ex_sizes[0] == (self.out_w, self.out_h)
            self.inpaint_mask = border_inpaint_mask(self.out_w, self.out_h, 0.17)
            self.inpaint_color = self.examples[0].copy()
        else:
            self.inpaint_mask = self.inpaint_color = None
        self.pyramids = [O.pyramid_build(e, levels) for e in self.examples]
        n = len(self.examples)
        self.method_list = list(self.methods) if self.methods is not None else [O.METHOD_ALL] * n
        self.mask_list = [None] * n
        if self.sample_masks:
            for i, (w, h) in enumerate(self.ex_sizes):
                if self.method_list[i] == O.METHOD_IMAGE:
                    mk = np.zeros((h, w, 4), np.uint8)
                    mk[..., 3] = 255
                    mk[h // 4:, : (3 * w) // 4, :3] = 255  # only this region may be sampled
                    self.mask_list[i] = mk
        self.guides = None
        if self.guided:
            tg = synth_texture(self.out_w, self.out_h, self.tex_seed + 100)
            tg[..., 1] = tg[..., 0]
            tg[..., 2] = tg[..., 0]
            exg = []
            for i, (w, h) in enumerate(self.guide_sizes or self.ex_sizes):
                gimg = synth_texture(w, h, self.tex_seed + 200 + i)
                gimg[..., 1] = gimg[..., 0]
                gimg[..., 2] = gimg[..., 0]
                exg.append(O.pyramid_build(gimg, levels))
            self.guides = (O.pyramid_build(tg, levels), exg)
        self._built = True
        return self

    def oracle_params(self, threads=1):
        return O.make_params(k=self.k, m=self.m, cauchy=self.cauchy, p=self.p, stages=self.stages, seed=self.seed,
                             alpha=self.alpha, threads=threads, tiling=self.tiling)

    def gpu_params(self):
        from texture_synthesis_b200 import capi
        return capi.make_params(k=self.k, m=self.m, cauchy=self.cauchy, p=self.p, stages=self.stages, seed=self.seed,
                                alpha=self.alpha, threads=1, tiling=self.tiling)

    # ---- oracle ------------------------------------------------------------------------------
    def oracle_generator(self, trace=False):
        self.build()
        g = O.Generator(self.out_w, self.out_h, self.inpaint_mask, self.inpaint_color, 0)
        g.set_examples(self.pyramids, self.method_list, self.mask_list)
        if self.guides is not None:
            g.set_guides(self.guides[0], self.guides[1])
        if trace:
            g.set_trace(True)
        if self.random_init:
            g.random_init(self.random_init, self.seed)
        return g

    def run_oracle(self, max_items=-1, trace=False, threads=1):
        g = self.oracle_generator(trace)
        g.resolve(self.oracle_params(threads), max_items)
        return g

    # ---- CUDA path ---------------------------------------------------------------------------
    def gpu_generator(self, trace=False, device=-1):
        from texture_synthesis_b200 import capi
        self.build()
        g = capi.Generator(self.out_w, self.out_h, self.inpaint_mask, self.inpaint_color, 0, device)
        if trace:
            g.set_trace(True)
        if self.random_init:
            g.random_init(self.random_init, [p[-1] for p in self.pyramids], self.seed)
        return g

    def run_gpu(self, trace=False):
        g = self.gpu_generator(trace)
        g.resolve(self.gpu_params(), self.pyramids, self.method_list, self.mask_list, self.guides)
        return g


def small_cases():
    M = O
    return [
        Case("single_64", 64, 64, [(48, 48)], seed=3),
        Case("single_rect", 100, 72, [(64, 56)], seed=120),
        Case("multi_randinit", 80, 80, [(40, 40), (48, 36), (32, 32)], seed=211, random_init=10),
        Case("tiling", 96, 96, [(64, 64)], seed=7, tiling=True),
        Case("inpaint_tiling", 96, 96, [(96, 96)], seed=5, tiling=True, inpaint=True),
        Case("inpaint", 80, 64, [(80, 64)], seed=9, inpaint=True),
        Case("masks_ignore", 72, 72, [(48, 48), (40, 40), (36, 36)], seed=211,
             methods=[M.METHOD_IMAGE, M.METHOD_IGNORE, M.METHOD_ALL], sample_masks=True),
        Case("guided", 72, 72, [(56, 56)], seed=2, guided=True),
        Case("k20_m10_s3", 64, 64, [(40, 40)], seed=1, k=20, m=10, stages=3, p=0.4, cauchy=0.7),
    ]


def compare_runs(go, gg, check_scores=True):
    """Returns a dict of mismatch statistics between an oracle generator and a CUDA generator."""
    co, cg = go.color(), gg.color()
    xo, xg = go.coord(), gg.coord()
    io, ig = go.ids(), gg.ids()
    fo, so = go.resolved()
    fg, sg = gg.resolved()
    out = dict(
        n=co.shape[0] * co.shape[1],
        color_mismatch=int((co != cg).any(axis=2).sum()),
        coord_mismatch=int((xo != xg).any(axis=2).sum()),
        id_mismatch=int((io != ig).any(axis=2).sum()),
        order_equal=bool(len(fo) == len(fg) and (fo == fg).all()),
    )
    if check_scores and len(so) == len(sg):
        denom = np.maximum(np.abs(so), 1e-30)
        out["score_max_rel"] = float(np.max(np.abs(so - sg) / denom)) if len(so) else 0.0
        out["score_bit_mismatch"] = int((so.view(np.uint32) != sg.view(np.uint32)).sum())
    return out


def edge_cases():
    """Ragged / extreme configurations (tiny outputs, 0 and 1 backtrack stages, k = 1, k + m at the limit, ...)."""
    return [
        Case("tiny_16", 16, 16, [(16, 16)], seed=1),
        Case("stages0", 40, 40, [(24, 24)], seed=2, stages=0),
        Case("stages1", 40, 40, [(24, 24)], seed=3, stages=1),
        Case("k5_m3", 48, 32, [(20, 28)], seed=4, k=5, m=3),
        Case("k1_m1", 24, 24, [(16, 16)], seed=5, k=1, m=1, stages=2),
        Case("long_37x91", 37, 91, [(30, 17)], seed=6),
        Case("randinit_many", 20, 20, [(16, 16)], seed=7, random_init=150),
        Case("p09", 48, 48, [(32, 32)], seed=8, p=0.9, stages=3),
        Case("k100_m100", 64, 64, [(40, 40)], seed=9, k=100, m=100, stages=3),
        Case("out_smaller_than_ex", 24, 24, [(96, 96)], seed=10),
        Case("tiling_tiny", 20, 20, [(16, 16)], seed=11, tiling=True),
        # q14: cauchy_dispersion = 0 passes the reference's validation (session.rs:451); x / 0 gives inf / NaN cost tables
        Case("cauchy0", 40, 40, [(24, 24)], seed=12, cauchy=0.0, stages=3),
        Case("cauchy0_multi", 36, 36, [(24, 24), (20, 28)], seed=13, cauchy=0.0, stages=2, k=12, m=9),
        # q13: mostly locked inpaint + guides drives adaptive_alpha negative (ms.rs:846-851), costs can be negative
        Case("neg_alpha", 64, 64, [(64, 64)], seed=14, inpaint=True, guided=True, stages=4),
        # an example guide smaller than its example keeps its own bounds (ms.rs:1265-1273)
        Case("guide_shorter", 64, 64, [(48, 48)], seed=15, guided=True, guide_sizes=[(48, 30)]),
        Case("guide_narrower", 64, 64, [(48, 48)], seed=16, guided=True, guide_sizes=[(31, 48)], stages=3),
    ]
