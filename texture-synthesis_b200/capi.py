"""ctypes binding of libtsb200.so -- the same C ABI (include/tsb200.h) the Rust shim binds.

The library is the product: if it is missing or no CUDA device is usable, calls fail loudly.
There is no CPU fallback anywhere in this package.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FILTER_TRIANGLE, FILTER_CATMULLROM, FILTER_GAUSSIAN = 0, 1, 2
SAMPLE_ALL, SAMPLE_IGNORE, SAMPLE_IMAGE = 0, 1, 2


class TsbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tsb200 error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    """tsb_params == GeneratorParams (reference lib/src/ms.rs:18-42)."""
    _fields_ = [
        ("nearest_neighbors", C.c_uint32), ("_pad0", C.c_uint32),
        ("random_sample_locations", C.c_uint64),
        ("cauchy_dispersion", C.c_float), ("p", C.c_float),
        ("p_stages", C.c_int32), ("alpha", C.c_float),
        ("seed", C.c_uint64), ("max_thread_count", C.c_uint64),
        ("tiling_mode", C.c_int32), ("_pad1", C.c_int32),
    ]


class Image(C.Structure):
    _fields_ = [("rgba", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32)]


class Pyramid(C.Structure):
    _fields_ = [("levels", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("n_levels", C.c_uint32)]


class Sampling(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("rgba", C.c_void_p)]


class Guides(C.Structure):
    _fields_ = [("target", Pyramid), ("examples", C.POINTER(Pyramid)), ("n_examples", C.c_uint32), ("_pad", C.c_uint32)]


class GeneratorDesc(C.Structure):
    _fields_ = [("out_width", C.c_uint32), ("out_height", C.c_uint32), ("inpaint_mask", C.c_void_p),
                ("inpaint_color", C.c_void_p), ("inpaint_example_index", C.c_uint32), ("device", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("work_items", C.c_uint64), ("candidates", C.c_uint64), ("texels_fetched", C.c_uint64),
                ("texels_nominal", C.c_uint64), ("rounds", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("phases", C.c_uint64), ("gpu_ms_resolve", C.c_double), ("gpu_ms_analysis", C.c_double),
                ("gpu_ms_other", C.c_double), ("host_ms_schedule", C.c_double), ("wall_ms_total", C.c_double),
                ("gpu_ms_total", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


BARRIER_FN = C.CFUNCTYPE(None, C.c_void_p)
PROGRESS_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64)

# every symbol include/tsb200.h declares
EXPORTS = [
    "tsb_pyramid_build", "tsb_resize", "tsb_generator_create", "tsb_generator_destroy", "tsb_generator_random_init",
    "tsb_generator_resolve", "tsb_generator_upload_inputs", "tsb_generator_resolve_resident", "tsb_generator_reset",
    "tsb_generator_read_color", "tsb_generator_read_coord", "tsb_generator_read_id", "tsb_generator_resolved_count",
    "tsb_generator_read_resolved", "tsb_generator_read_uncertainty", "tsb_generator_read_id_maps",
    "tsb_generator_get_stats", "tsb_last_error", "tsb_device_count", "tsb_generator_load_state",
    "tsb_generator_eval_items", "tsb_generator_set_trace", "tsb_generator_trace_count", "tsb_generator_read_trace",
    "tsb_microbench_gather", "tsb_generator_mg_prepare", "tsb_generator_mg_export", "tsb_generator_mg_attach",
    "tsb_generator_mg_phases", "tsb_guide_map", "tsb_match_histograms",
]


def library_path():
    return os.path.join(_HERE, "libtsb200.so")


def lib():
    global _LIB
    if _LIB is None:
        so = library_path()
        if not os.path.exists(so):
            raise ImportError(f"{so} is missing: build it with `python texture-synthesis_b200/build.py` "
                              "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(so)
        L.tsb_last_error.restype = C.c_char_p
        vp = C.c_void_p
        L.tsb_pyramid_build.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, vp]
        L.tsb_resize.argtypes = [vp, C.c_uint32, C.c_uint32, vp, C.c_uint32, C.c_uint32, C.c_int]
        L.tsb_generator_create.argtypes = [C.POINTER(GeneratorDesc), C.POINTER(vp)]
        L.tsb_guide_map.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_float, vp]
        L.tsb_match_histograms.argtypes = [vp, C.c_uint32, C.c_uint32, vp, C.c_uint32, C.c_uint32, vp]
        L.tsb_generator_destroy.argtypes = [vp]
        L.tsb_generator_destroy.restype = None
        L.tsb_generator_random_init.argtypes = [vp, C.c_uint64, C.POINTER(Image), C.c_uint32, C.c_uint64]
        L.tsb_generator_resolve.argtypes = [vp, C.POINTER(Params), C.POINTER(Pyramid), C.c_uint32, C.POINTER(Guides),
                                            C.POINTER(Sampling), PROGRESS_FN, vp]
        L.tsb_generator_upload_inputs.argtypes = [vp, C.POINTER(Pyramid), C.c_uint32, C.POINTER(Guides), C.POINTER(Sampling)]
        L.tsb_generator_resolve_resident.argtypes = [vp, C.POINTER(Params), PROGRESS_FN, vp]
        L.tsb_generator_reset.argtypes = [vp]
        for n in ("tsb_generator_read_color", "tsb_generator_read_coord", "tsb_generator_read_id", "tsb_generator_read_uncertainty"):
            getattr(L, n).argtypes = [vp, vp]
        L.tsb_generator_read_id_maps.argtypes = [vp, vp, vp]
        L.tsb_generator_resolved_count.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.tsb_generator_read_resolved.argtypes = [vp, vp, vp]
        L.tsb_generator_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.tsb_generator_load_state.argtypes = [vp, vp, vp, vp, vp, C.c_uint64, vp, vp, C.c_uint64, C.c_uint64]
        L.tsb_generator_eval_items.argtypes = [vp, C.POINTER(Params), C.c_int32, C.c_float, C.c_uint64, C.c_uint32,
                                               vp, vp, vp, vp, vp]
        L.tsb_generator_set_trace.argtypes = [vp, C.c_int]
        L.tsb_generator_trace_count.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.tsb_generator_read_trace.argtypes = [vp, vp, vp, vp, vp, vp]
        L.tsb_microbench_gather.argtypes = [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.tsb_generator_mg_prepare.argtypes = [vp, C.POINTER(Params), C.POINTER(C.c_uint32)]
        L.tsb_generator_mg_export.argtypes = [vp, vp]
        L.tsb_generator_mg_attach.argtypes = [vp, C.c_uint32, C.c_uint32, vp, BARRIER_FN, vp]
        L.tsb_generator_mg_phases.argtypes = [vp, C.POINTER(C.c_uint64)]
        _LIB = L
    return _LIB


def _check(rc):
    if rc != 0:
        raise TsbError(rc, lib().tsb_last_error().decode("utf-8", "replace"))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _rgba(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if a.ndim != 3 or a.shape[2] != 4:
        raise ValueError("expected an RGBA8 array [h, w, 4]")
    return a


def device_count():
    return lib().tsb_device_count()


def make_params(k=50, m=50, cauchy=1.0, p=0.5, stages=5, seed=0, alpha=0.8, threads=1, tiling=False):
    """Defaults of reference lib/src/lib.rs:343-359."""
    return Params(k, 0, m, cauchy, p, stages, alpha, seed, threads, 1 if tiling else 0, 0)


def resize(img, nw, nh, filt):
    img = _rgba(img)
    h, w = img.shape[:2]
    out = np.empty((nh, nw, 4), np.uint8)
    _check(lib().tsb_resize(_p(img), w, h, _p(out), nw, nh, filt))
    return out


def pyramid_build(img, levels):
    """ImagePyramid::build_gaussian (reference lib/src/img_pyramid.rs:20-37) -> [levels, h, w, 4]."""
    img = _rgba(img)
    h, w = img.shape[:2]
    out = np.empty((max(1, levels), h, w, 4), np.uint8)
    _check(lib().tsb_pyramid_build(_p(img), w, h, levels, _p(out)))
    return out


def guide_map(img, sigma=2.0):
    """utils::transform_to_guide_map (reference lib/src/utils.rs:101-116) on the GPU."""
    img = _rgba(img)
    h, w = img.shape[:2]
    out = np.empty((h, w, 4), np.uint8)
    _check(lib().tsb_guide_map(_p(img), w, h, sigma, _p(out)))
    return out


def match_histograms(source, target):
    """utils::match_histograms (reference lib/src/utils.rs:135-183) on the GPU; returns the remapped source."""
    source, target = _rgba(source), _rgba(target)
    out = np.empty_like(source)
    _check(lib().tsb_match_histograms(_p(source), source.shape[1], source.shape[0], _p(target), target.shape[1], target.shape[0], _p(out)))
    return out


def microbench_gather(nbytes, mode=0, iters=20):
    gbs, gps = C.c_double(), C.c_double()
    _check(lib().tsb_microbench_gather(nbytes, mode, iters, C.byref(gbs), C.byref(gps)))
    return gbs.value, gps.value


def _pyr_struct(p):
    return Pyramid(p.ctypes.data, p.shape[2], p.shape[1], p.shape[0])


class Generator:
    """Handle on one device-side Generator (reference lib/src/ms.rs:207-217)."""

    def __init__(self, out_w, out_h, inpaint_mask=None, inpaint_color=None, inpaint_index=0, device=-1):
        self.L = lib()
        self.W, self.H = int(out_w), int(out_h)
        self._keep = []
        if inpaint_mask is not None:
            inpaint_mask, inpaint_color = _rgba(inpaint_mask), _rgba(inpaint_color)
            if inpaint_mask.shape != (self.H, self.W, 4) or inpaint_color.shape != (self.H, self.W, 4):
                raise ValueError("inpaint mask / colour must have the output size")
        desc = GeneratorDesc(self.W, self.H, _p(inpaint_mask), _p(inpaint_color), inpaint_index, device)
        h = C.c_void_p()
        _check(self.L.tsb_generator_create(C.byref(desc), C.byref(h)))
        self.h = h
        self.k = 50

    def close(self):
        if getattr(self, "h", None):
            self.L.tsb_generator_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    # -- inputs -------------------------------------------------------------------------------
    def _marshal_inputs(self, pyramids, methods, masks, guides):
        pyramids = [np.ascontiguousarray(p, np.uint8) for p in pyramids]
        n = len(pyramids)
        parr = (Pyramid * n)(*[_pyr_struct(p) for p in pyramids])
        methods = list(methods) if methods is not None else [SAMPLE_ALL] * n
        mk = [None if (masks is None or masks[i] is None) else _rgba(masks[i]) for i in range(n)]
        sarr = (Sampling * n)(*[Sampling(methods[i], 0, mk[i].ctypes.data if mk[i] is not None else None) for i in range(n)])
        gptr, keep = None, [pyramids, parr, mk, sarr]
        if guides is not None:
            target, exg = guides
            target = np.ascontiguousarray(target, np.uint8)
            exg = [np.ascontiguousarray(g, np.uint8) for g in exg]
            garr = (Pyramid * len(exg))(*[_pyr_struct(g) for g in exg])
            gs = Guides(_pyr_struct(target), garr, len(exg), 0)
            gptr = C.pointer(gs)
            keep += [target, exg, garr, gs]
        return n, parr, sarr, gptr, keep

    def upload_inputs(self, pyramids, methods=None, masks=None, guides=None):
        n, parr, sarr, gptr, keep = self._marshal_inputs(pyramids, methods, masks, guides)
        _check(self.L.tsb_generator_upload_inputs(self.h, parr, n, gptr, sarr))

    def random_init(self, count, top_levels, seed):
        imgs = [_rgba(t) for t in top_levels]
        arr = (Image * len(imgs))(*[Image(i.ctypes.data, i.shape[1], i.shape[0]) for i in imgs])
        _check(self.L.tsb_generator_random_init(self.h, count, arr, len(imgs), seed))

    # -- hot path -----------------------------------------------------------------------------
    def _cb(self, progress):
        if progress is None:
            return C.cast(None, PROGRESS_FN)

        def tramp(user, rgba, w, h, tc, tt, sc, st):
            img = np.ctypeslib.as_array(C.cast(rgba, C.POINTER(C.c_uint8)), shape=(h, w, 4))
            progress(img, (tc, tt), (sc, st))
        return PROGRESS_FN(tramp)

    def resolve(self, params, pyramids, methods=None, masks=None, guides=None, progress=None):
        """tsb_generator_resolve: host buffers in, blocking (Generator::resolve, ms.rs:702)."""
        n, parr, sarr, gptr, keep = self._marshal_inputs(pyramids, methods, masks, guides)
        self.k = params.nearest_neighbors
        cb = self._cb(progress)
        _check(self.L.tsb_generator_resolve(self.h, C.byref(params), parr, n, gptr, sarr, cb, None))

    def resolve_resident(self, params, progress=None):
        self.k = params.nearest_neighbors
        cb = self._cb(progress)
        _check(self.L.tsb_generator_resolve_resident(self.h, C.byref(params), cb, None))

    def reset(self):
        _check(self.L.tsb_generator_reset(self.h))

    # -- read-outs ----------------------------------------------------------------------------
    def color(self):
        out = np.empty((self.H, self.W, 4), np.uint8)
        _check(self.L.tsb_generator_read_color(self.h, _p(out)))
        return out

    def coord(self):
        out = np.empty((self.H, self.W, 3), np.uint32)
        _check(self.L.tsb_generator_read_coord(self.h, _p(out)))
        return out

    def ids(self):
        out = np.empty((self.H, self.W, 2), np.uint32)
        _check(self.L.tsb_generator_read_id(self.h, _p(out)))
        return out

    def resolved(self):
        n, locked = C.c_uint64(), C.c_uint64()
        _check(self.L.tsb_generator_resolved_count(self.h, C.byref(n), C.byref(locked)))
        flat = np.empty(n.value, np.uint32)
        score = np.empty(n.value, np.float32)
        _check(self.L.tsb_generator_read_resolved(self.h, _p(flat), _p(score)))
        return flat, score

    def locked_count(self):
        n, locked = C.c_uint64(), C.c_uint64()
        _check(self.L.tsb_generator_resolved_count(self.h, C.byref(n), C.byref(locked)))
        return locked.value

    def uncertainty_map(self):
        out = np.empty((self.H, self.W, 4), np.uint8)
        _check(self.L.tsb_generator_read_uncertainty(self.h, _p(out)))
        return out

    def id_maps(self):
        a = np.empty((self.H, self.W, 4), np.uint8)
        b = np.empty((self.H, self.W, 4), np.uint8)
        _check(self.L.tsb_generator_read_id_maps(self.h, _p(a), _p(b)))
        return a, b

    def stats(self):
        s = Stats()
        _check(self.L.tsb_generator_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    # -- parity harness -----------------------------------------------------------------------
    def load_state(self, color, coord, ids, tree_xy, resolved_flat, resolved_score, locked):
        color = _rgba(color)
        coord = np.ascontiguousarray(coord, np.uint32)
        ids = np.ascontiguousarray(ids, np.uint32)
        tree_xy = np.ascontiguousarray(tree_xy, np.int32)
        rf = np.ascontiguousarray(resolved_flat, np.uint32)
        rs = np.ascontiguousarray(resolved_score, np.float32)
        _check(self.L.tsb_generator_load_state(self.h, _p(color), _p(coord), _p(ids), _p(tree_xy), len(tree_xy),
                                               _p(rf), _p(rs), len(rf), locked))

    def eval_items(self, params, level, adaptive_alpha, p_stage_seed, pixels, loop_seeds):
        pixels = np.ascontiguousarray(pixels, np.uint32)
        loop_seeds = np.ascontiguousarray(loop_seeds, np.uint64)
        n, k = len(pixels), params.nearest_neighbors
        neigh = np.empty((n, k, 2), np.int32)
        res = np.empty((n, 8), np.int32)
        score = np.empty(n, np.float32)
        _check(self.L.tsb_generator_eval_items(self.h, C.byref(params), level, adaptive_alpha, p_stage_seed, n,
                                               _p(pixels), _p(loop_seeds), _p(neigh), _p(res), _p(score)))
        return dict(neigh=neigh, res=res, score=score)

    # -- band-sharded multi-GPU ----------------------------------------------------------------
    def mg_export(self, params):
        """Allocates the shared buffers and returns this rank's CUDA IPC handles as bytes."""
        n = C.c_uint32()
        _check(self.L.tsb_generator_mg_prepare(self.h, C.byref(params), C.byref(n)))
        buf = np.zeros(n.value * 80, np.uint8)
        _check(self.L.tsb_generator_mg_export(self.h, _p(buf)))
        return buf.tobytes()

    def mg_attach(self, rank, world, all_handles, barrier):
        """all_handles: list (by rank) of the bytes returned by mg_export; barrier: callable blocking until all ranks arrive."""
        blob = np.frombuffer(b"".join(all_handles), np.uint8).copy()
        self._barrier_cb = BARRIER_FN(lambda user: barrier())
        self._keep.append(blob)
        _check(self.L.tsb_generator_mg_attach(self.h, rank, world, _p(blob), self._barrier_cb, None))

    def mg_phases(self):
        n = C.c_uint64()
        _check(self.L.tsb_generator_mg_phases(self.h, C.byref(n)))
        return n.value

    def set_trace(self, on=True):
        _check(self.L.tsb_generator_set_trace(self.h, 1 if on else 0))

    def trace(self):
        n = C.c_uint64()
        _check(self.L.tsb_generator_trace_count(self.h, C.byref(n)))
        n = n.value
        px = np.empty(n, np.uint32)
        best = np.empty(n, np.int32)
        ncand = np.empty(n, np.int32)
        nneigh = np.empty(n, np.int32)
        score = np.empty(n, np.float32)
        _check(self.L.tsb_generator_read_trace(self.h, _p(px), _p(best), _p(ncand), _p(nneigh), _p(score)))
        return dict(pixel=px, best=best, ncand=ncand, nneigh=nneigh, score=score)
