"""CPU tests: the C-ABI library loads, exports every symbol the header declares, and fails loudly without a GPU."""
import ctypes as C
import io
import os
import re

import numpy as np
import pytest

import texture_synthesis_b200 as ts
from texture_synthesis_b200 import capi
from tests.helpers import gpu_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "tsb200.h")).read()
    declared = set(re.findall(r"\b(tsb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"tsb_progress_fn"}
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)
    L = capi.lib()
    for name in declared:
        assert getattr(L, name) is not None


def test_struct_layouts_match_header():
    assert C.sizeof(capi.Params) == 56 and capi.Params.seed.offset == 32 and capi.Params.tiling_mode.offset == 48
    assert C.sizeof(capi.Pyramid) == 24 and C.sizeof(capi.Image) == 16 and C.sizeof(capi.Sampling) == 16
    assert C.sizeof(capi.GeneratorDesc) == 32 and C.sizeof(capi.Stats) == 104


@pytest.mark.skipif(gpu_available(), reason="checks the no-GPU failure path")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(capi.TsbError) as e:
        capi.Generator(32, 32)
    assert e.value.code == -2                                            # TSB_ERR_CUDA
    with pytest.raises(capi.TsbError):
        capi.pyramid_build(np.zeros((16, 16, 4), np.uint8), 3)


def test_session_parameter_validation_mirrors_reference():
    img = np.zeros((16, 16, 4), np.uint8)
    for setter, bad, name in (("cauchy_dispersion", 1.5, "cauchy-dispersion"), ("backtrack_percent", -0.1, "backtrack-percent"),
                              ("guide_alpha", 2.0, "guide-alpha"), ("max_thread_count", 0, "max-thread-count"),
                              ("random_sample_locations", 0, "m-rand")):
        b = ts.Session.builder().add_example(img)
        getattr(b, setter)(bad)
        with pytest.raises(ts.InvalidRange) as e:
            b.build()
        assert e.value.name == name
    with pytest.raises(ts.Error) as e:
        ts.Session.builder().build()
    assert e.value.kind == "NoExamples"
    with pytest.raises(ts.Error) as e:
        ts.Session.builder().add_example(ts.Example(img).set_sample_method(ts.SampleMethod.Ignore())).build()
    assert e.value.kind == "NoExamples"
    with pytest.raises(ts.Error) as e:
        ts.Session.builder().add_example(ts.Example(img).with_guide(img)).add_example(img).build()
    assert e.value.kind == "ExampleGuideMismatch"


def test_coordinate_transform_serde_roundtrip():
    # mirrors the reference's only unit test, coord_tx_serde (lib/src/lib.rs:642-693)
    rng = np.random.RandomState(0)
    buf = rng.randint(0, 400, size=(30, 20, 3)).astype(np.uint32)
    ct = ts.CoordinateTransform(buf, ts.Dims(20, 30), [ts.Dims(400, 300), ts.Dims(64, 64)])
    f = io.BytesIO()
    ct.write(f)
    raw = f.getvalue()
    assert np.frombuffer(raw[:16], np.uint32).tolist() == [0x12340001, 20, 30, 2]
    assert len(raw) == 4 * (4 + 4 + 20 * 30 * 3)
    back = ts.CoordinateTransform.read(io.BytesIO(raw))
    assert back.output_size == ct.output_size and back.original_maps == ct.original_maps and (back.buffer == buf).all()
    with pytest.raises(ts.Error):
        ts.CoordinateTransform.read(io.BytesIO(b"\x00" * 64))
