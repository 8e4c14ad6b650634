"""ctypes binding of the CPU oracle (oracle/libts_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

F_TRIANGLE, F_CATMULLROM, F_GAUSSIAN = 0, 1, 2
METHOD_ALL, METHOD_IGNORE, METHOD_IMAGE = 0, 1, 2


class Params(C.Structure):
    """GeneratorParams (ms.rs:18-42); identical layout to tsb_params in include/tsb200.h."""
    _fields_ = [
        ("nearest_neighbors", C.c_uint32), ("_pad0", C.c_uint32),
        ("random_sample_locations", C.c_uint64),
        ("cauchy_dispersion", C.c_float), ("p", C.c_float),
        ("p_stages", C.c_int32), ("alpha", C.c_float),
        ("seed", C.c_uint64), ("max_thread_count", C.c_uint64),
        ("tiling_mode", C.c_int32), ("_pad1", C.c_int32),
    ]


def make_params(k=50, m=50, cauchy=1.0, p=0.5, stages=5, seed=0, alpha=0.8, threads=1, tiling=False):
    """Defaults of lib.rs:343-359."""
    return Params(k, 0, m, cauchy, p, stages, alpha, seed, threads, 1 if tiling else 0, 0)


def build(force=False):
    so = os.path.join(_HERE, "libts_oracle.so")
    src = os.path.join(_HERE, "ts_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libts_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_gen_create.restype = C.c_void_p
        L.orc_gen_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_gen_destroy.argtypes = [C.c_void_p]
        L.orc_gen_set_examples.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_gen_set_guides.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_gen_random_init.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.orc_gen_set_trace.argtypes = [C.c_void_p, C.c_int]
        L.orc_gen_resolve.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int64]
        L.orc_gen_last_seconds.restype = C.c_double
        L.orc_gen_last_seconds.argtypes = [C.c_void_p]
        for n in ("orc_gen_read_color", "orc_gen_read_coord", "orc_gen_read_id", "orc_gen_read_tree_points",
                  "orc_uncertainty_map"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_void_p]
        L.orc_id_maps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        for n in ("orc_gen_resolved_count", "orc_gen_locked_count", "orc_gen_tree_point_count", "orc_gen_trace_count"):
            getattr(L, n).restype = C.c_uint64
            getattr(L, n).argtypes = [C.c_void_p]
        L.orc_gen_read_resolved.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_gen_read_trace.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.orc_gen_eval_items.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int, C.c_float, C.c_uint64, C.c_uint32,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_pcg32_new_stream.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
        L.orc_pcg32_from_seed_next_u64.restype = C.c_uint64
        L.orc_pcg32_from_seed_next_u64.argtypes = [C.c_void_p]
        L.orc_pcg32_seed_from_u64.argtypes = [C.c_uint64, C.c_uint32, C.c_void_p]
        L.orc_gen_range_seq.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_uint32, C.c_void_p]
        L.orc_resize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_pyramid_build.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_void_p]
        L.orc_guide_map.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
        L.orc_match_histograms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def resize(img, nw, nh, filt):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    out = np.empty((nh, nw, 4), np.uint8)
    lib().orc_resize(_p(img), w, h, _p(out), nw, nh, filt)
    return out


def pyramid_build(img, levels):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    n = max(1, levels)
    out = np.empty((n, h, w, 4), np.uint8)
    lib().orc_pyramid_build(_p(img), w, h, levels, _p(out))
    return out


def guide_map(img, sigma=2.0):
    """transform_to_guide_map (utils.rs:101-116)"""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    out = np.empty((h, w, 4), np.uint8)
    lib().orc_guide_map(_p(img), w, h, sigma, _p(out))
    return out


def match_histograms(source, target):
    """match_histograms (utils.rs:135-163); returns the modified copy of source"""
    src = np.ascontiguousarray(source, dtype=np.uint8).copy()
    tgt = np.ascontiguousarray(target, dtype=np.uint8)
    lib().orc_match_histograms(_p(src), src.shape[1], src.shape[0], _p(tgt), tgt.shape[1], tgt.shape[0])
    return src


class Generator:
    """Oracle-side Generator (ms.rs:207-217) driven the way Session::build/run drive it."""

    def __init__(self, out_w, out_h, inpaint_mask=None, inpaint_color=None, inpaint_index=0):
        self.L = lib()
        self.W, self.H = out_w, out_h
        self._keep = []
        if inpaint_mask is not None:
            inpaint_mask = np.ascontiguousarray(inpaint_mask, np.uint8)
            inpaint_color = np.ascontiguousarray(inpaint_color, np.uint8)
            assert inpaint_mask.shape == (out_h, out_w, 4) and inpaint_color.shape == (out_h, out_w, 4)
            self._keep += [inpaint_mask, inpaint_color]
        self.h = self.L.orc_gen_create(out_w, out_h, _p(inpaint_mask), _p(inpaint_color), inpaint_index)
        self.k = 50

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_gen_destroy(self.h)
            self.h = None

    def set_examples(self, pyramids, methods=None, masks=None):
        """pyramids: list of uint8 arrays [levels, h, w, 4] (level 0 = blurriest)."""
        n = len(pyramids)
        pyramids = [np.ascontiguousarray(p, np.uint8) for p in pyramids]
        levels = pyramids[0].shape[0]
        ws = np.array([p.shape[2] for p in pyramids], np.int32)
        hs = np.array([p.shape[1] for p in pyramids], np.int32)
        ptrs = (C.c_void_p * n)(*[p.ctypes.data for p in pyramids])
        methods = np.array(methods if methods is not None else [METHOD_ALL] * n, np.int32)
        mk = [None if (masks is None or masks[i] is None) else np.ascontiguousarray(masks[i], np.uint8) for i in range(n)]
        mptrs = (C.c_void_p * n)(*[(m.ctypes.data if m is not None else None) for m in mk])
        self._keep += [pyramids, ws, hs, ptrs, methods, mk, mptrs]
        self.L.orc_gen_set_examples(self.h, n, levels, _p(ws), _p(hs), ptrs, _p(methods), mptrs)
        self.levels = levels

    def set_guides(self, target_pyr, example_guide_pyrs):
        target_pyr = np.ascontiguousarray(target_pyr, np.uint8)
        gs = [np.ascontiguousarray(g, np.uint8) for g in example_guide_pyrs]
        n = len(gs)
        gw = np.array([g.shape[2] for g in gs], np.int32)
        gh = np.array([g.shape[1] for g in gs], np.int32)
        ptrs = (C.c_void_p * n)(*[g.ctypes.data for g in gs])
        self._keep += [target_pyr, gs, gw, gh, ptrs]
        self.L.orc_gen_set_guides(self.h, _p(target_pyr), target_pyr.shape[2], target_pyr.shape[1], n, _p(gw), _p(gh), ptrs)

    def random_init(self, count, seed):
        self.L.orc_gen_random_init(self.h, count, seed)

    def set_trace(self, on=True):
        self.L.orc_gen_set_trace(self.h, 1 if on else 0)

    def resolve(self, params, max_items=-1):
        self.k = params.nearest_neighbors
        self.L.orc_gen_resolve(self.h, C.byref(params), max_items)
        return self.L.orc_gen_last_seconds(self.h)

    def color(self):
        out = np.empty((self.H, self.W, 4), np.uint8)
        self.L.orc_gen_read_color(self.h, _p(out))
        return out

    def coord(self):
        out = np.empty((self.H, self.W, 3), np.uint32)
        self.L.orc_gen_read_coord(self.h, _p(out))
        return out

    def ids(self):
        out = np.empty((self.H, self.W, 2), np.uint32)
        self.L.orc_gen_read_id(self.h, _p(out))
        return out

    def resolved(self):
        n = self.L.orc_gen_resolved_count(self.h)
        flat = np.empty(n, np.uint32)
        score = np.empty(n, np.float32)
        self.L.orc_gen_read_resolved(self.h, _p(flat), _p(score))
        return flat, score

    def locked_count(self):
        return self.L.orc_gen_locked_count(self.h)

    def tree_points(self):
        n = self.L.orc_gen_tree_point_count(self.h)
        xy = np.empty((n, 2), np.int32)
        self.L.orc_gen_read_tree_points(self.h, _p(xy))
        return xy

    def trace(self):
        n = self.L.orc_gen_trace_count(self.h)
        px = np.empty(n, np.uint32)
        best = np.empty(n, np.int32)
        ncand = np.empty(n, np.int32)
        nneigh = np.empty(n, np.int32)
        score = np.empty(n, np.float32)
        self.L.orc_gen_read_trace(self.h, _p(px), _p(best), _p(ncand), _p(nneigh), _p(score))
        return dict(pixel=px, best=best, ncand=ncand, nneigh=nneigh, score=score)

    def eval_items(self, params, level, adaptive_alpha, p_stage_seed, pixels, loop_seeds):
        pixels = np.ascontiguousarray(pixels, np.uint32)
        loop_seeds = np.ascontiguousarray(loop_seeds, np.uint64)
        n = len(pixels)
        k = params.nearest_neighbors
        neigh = np.empty((n, k, 2), np.int32)
        res = np.empty((n, 8), np.int32)
        score = np.empty(n, np.float32)
        self.L.orc_gen_eval_items(self.h, C.byref(params), level, adaptive_alpha, p_stage_seed, n,
                                  _p(pixels), _p(loop_seeds), _p(neigh), _p(res), _p(score))
        return dict(neigh=neigh, res=res, score=score)

    def uncertainty_map(self):
        out = np.empty((self.H, self.W, 4), np.uint8)
        self.L.orc_uncertainty_map(self.h, _p(out))
        return out

    def id_maps(self):
        a = np.empty((self.H, self.W, 4), np.uint8)
        b = np.empty((self.H, self.W, 4), np.uint8)
        self.L.orc_id_maps(self.h, _p(a), _p(b))
        return a, b


def pcg32_new_stream(state, stream, n):
    out = np.empty(n, np.uint32)
    lib().orc_pcg32_new_stream(state, stream, n, _p(out))
    return out


def pcg32_from_seed_next_u64(seed16):
    s = np.ascontiguousarray(seed16, np.uint8)
    return lib().orc_pcg32_from_seed_next_u64(_p(s))


def pcg32_seed_from_u64(seed, n):
    out = np.empty(n, np.uint32)
    lib().orc_pcg32_seed_from_u64(seed, n, _p(out))
    return out


def gen_range_seq(seed, kind, n, count):
    out = np.empty(count, np.uint64)
    lib().orc_gen_range_seq(seed, kind, n, count, _p(out))
    return out
