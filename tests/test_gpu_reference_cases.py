"""The reference's own integration tests (lib/tests/diff.rs:163-252) on the CUDA path, through the Session mirror.

The reference pins each configuration with an 8x8 perceptual hash of its output (tolerant to per-pixel
differences, and not computable offline); here the same nine configurations are run on frozen decoded crops of
the same images (tests/golden/ref_imgs.npz, made by tests/golden/make_ref_inputs.py) and compared PIXEL FOR PIXEL
with the oracle's max_thread_count(1) pipeline."""
import os

import numpy as np
import pytest

import texture_synthesis_b200 as ts
from tests.helpers import O
from tests.oracle_session import OracleSession

pytestmark = pytest.mark.gpu
IMGS = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_imgs.npz"))
D100 = ts.Dims.square(100)


def _check(gpu_builder, oracle_session):
    got = gpu_builder.max_thread_count(1).build().run(None)
    want = oracle_session.run()
    assert (got.into_image() == want.color()).all()
    assert (got.get_coordinate_transform().buffer == want.coord()).all()
    return got, want


def test_single_example():
    o = OracleSession().add_example(IMGS["img1"])
    o.p["seed"], o.out_size = 120, (100, 100)
    _check(ts.Session.builder().add_example(IMGS["img1"]).seed(120).output_size(D100), o)


def test_multi_example():
    names = ["multi1", "multi2", "multi3", "multi4"]
    o = OracleSession()
    for n in names:
        o.add_example(IMGS[n])
    o.resize, o.random_init_count, o.out_size = (100, 100), 10, (100, 100)
    o.p["seed"] = 211
    got, want = _check(ts.Session.builder().add_examples([IMGS[n] for n in names]).resize_input(D100).random_init(10).seed(211).output_size(D100), o)
    pm_o, mm_o = want.id_maps()                         # save_debug's map_id.png (config 2 of BASELINE.json)
    pm_g, mm_g = got.inner.id_maps()
    assert (mm_o == mm_g).all() and (pm_o == pm_g).all()
    assert len(np.unique(mm_g.reshape(-1, 4), axis=0)) > 1


def test_guided():
    o = OracleSession().add_example(IMGS["img2"], guide=IMGS["mask_2_example"])
    o.target_guide, o.out_size = IMGS["mask_2_target"], (100, 100)
    _check(ts.Session.builder().add_example(ts.Example(IMGS["img2"]).with_guide(IMGS["mask_2_example"]))
           .load_target_guide(IMGS["mask_2_target"]).output_size(D100), o)


def test_style_transfer():
    o = OracleSession().add_example(IMGS["multi4"])
    o.target_guide, o.out_size = IMGS["tom"], (100, 100)
    _check(ts.Session.builder().add_example(IMGS["multi4"]).load_target_guide(IMGS["tom"]).output_size(D100), o)


def test_inpaint():
    o = OracleSession().inpaint_example(IMGS["mask_3_inpaint"], IMGS["img3"], (100, 100), method=O.METHOD_IMAGE, sample_mask=IMGS["mask_3_inpaint"])
    _check(ts.Session.builder().inpaint_example(IMGS["mask_3_inpaint"], ts.Example(IMGS["img3"]).set_sample_method(IMGS["mask_3_inpaint"]), D100), o)


def test_inpaint_channel():
    o = OracleSession().inpaint_example_channel("A", IMGS["bricks"], (120, 120))
    _check(ts.Session.builder().inpaint_example_channel("A", IMGS["bricks"], ts.Dims.square(120)), o)


def test_tiling():
    o = OracleSession().inpaint_example(IMGS["mask_1_tile"], IMGS["img1"], (100, 100))
    o.p["tiling"] = True
    _check(ts.Session.builder().inpaint_example(IMGS["mask_1_tile"], ts.Example(IMGS["img1"]), D100).tiling_mode(True), o)


def test_sample_masks():
    o = OracleSession().add_example(IMGS["img4"], method=O.METHOD_IMAGE, mask=IMGS["mask_4_sample"])
    o.p["seed"], o.out_size = 211, (100, 100)
    _check(ts.Session.builder().add_example(ts.Example(IMGS["img4"]).set_sample_method(IMGS["mask_4_sample"])).seed(211).output_size(D100), o)


def test_sample_masks_ignore():
    o = OracleSession().add_example(IMGS["img4"], method=O.METHOD_IGNORE).add_example(IMGS["img5"])
    o.p["seed"], o.out_size = 211, (120, 120)
    _check(ts.Session.builder().add_example(ts.Example(IMGS["img4"]).set_sample_method(ts.SampleMethod.Ignore()))
           .add_example(ts.Example(IMGS["img5"]).set_sample_method(ts.SampleMethod.All())).seed(211).output_size(ts.Dims.square(120)), o)
