"""The BASELINE.json configurations at their STATED sizes, and the reference's nine integration configurations
(lib/tests/diff.rs:163-252) at the reference's real sizes, as declarative specs that both the CPU oracle pipeline
(tests/oracle_session.py) and the CUDA Session mirror (texture_synthesis_b200.session) consume.

Inputs: the reference's own images decoded ONCE in the build container (tests/golden/ref_imgs_full.npz, made by
tests/golden/make_ref_inputs.py; the GPU box has no /root/reference) and the deterministic synthetic textures of
texture_synthesis_b200.synth.  The oracle's results are committed as SHA-256 digests
(tests/golden/fullsize_digests.json, made by tests/golden/make_fullsize_digests.py, max_thread_count(1) as
lib/tests/diff.rs:143-145 forces); tests/test_gpu_fullsize.py compares the CUDA path against them.
"""
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DIGESTS = os.path.join(HERE, "golden", "fullsize_digests.json")
_IMGS = {}
_DECODER = "pillow"


def set_decoder(name):
    """"pillow": the libjpeg decodes every digest and every CUDA parity test is built on (default).  "jpegport": the same
    images with the JPEG-derived ones as the restated jpeg-decoder 0.1.22 delivers them (oracle/jpeg_port.py; stored as
    differences in tests/golden/ref_imgs_jpegport.npz) -- used by tests/test_oracle_pin.py only."""
    global _DECODER
    assert name in ("pillow", "jpegport")
    _DECODER = name


def imgs():
    if _DECODER not in _IMGS:
        z = np.load(os.path.join(HERE, "golden", "ref_imgs_full.npz"))
        delta = np.load(os.path.join(HERE, "golden", "ref_imgs_jpegport.npz")) if _DECODER == "jpegport" else None
        d = {}
        for k in z.files:
            a = z[k]
            if delta is not None and k in delta.files:
                a = a.copy()
                a[..., :3] = (a[..., :3].astype(np.int16) + delta[k]).astype(np.uint8)
            if a.shape[2] == 3:  # alpha == 255 everywhere: stored as RGB
                a = np.concatenate([a, np.full(a.shape[:2] + (1,), 255, np.uint8)], axis=-1)
            d[k] = np.ascontiguousarray(a, np.uint8)
        _IMGS[_DECODER] = d
    return _IMGS[_DECODER]


def _synth(w, h, seed):
    from texture_synthesis_b200.synth import synth_texture
    return synth_texture(w, h, seed)


def _border_mask(w, h):
    from texture_synthesis_b200.synth import border_inpaint_mask
    return border_inpaint_mask(w, h, 0.17)


def ex(img, guide=None, method="all", mask=None):
    return dict(img=img, guide=guide, method=method, mask=mask)


# name -> function returning the spec (built lazily: the large synthetic inputs take a moment)
def _specs():
    I = imgs
    S = {}
    # ---- BASELINE.json configs at their stated sizes -------------------------------------------------------------
    # headline metric: 2048^2 from a synthetic 512^2 example, defaults
    S["headline_2048_from_512"] = lambda: dict(examples=[ex(_synth(512, 512, 1))], out=(2048, 2048))
    # C1: lib/examples/01_single_example_synthesis.rs:5-8 (imgs/1.jpg, defaults -> 500^2, seed 0)
    S["c1_single_example_500"] = lambda: dict(examples=[ex(I()["img1"])])
    # C2: lib/examples/02_multi_example_synthesis.rs:7-23 (4 examples resized to 300^2, random_init 10, seed 211) + map_id
    S["c2_multi_example_500"] = lambda: dict(examples=[ex(I()[f"multi{i}"]) for i in (1, 2, 3, 4)], resize=(300, 300),
                                             random_init=10, seed=211, id_maps=True)
    # C3: lib/examples/03_guided_synthesis.rs:4-11 and 04_style_transfer.rs:4-11
    S["c3_guided_500"] = lambda: dict(examples=[ex(I()["img2"], guide=I()["mask_2_example"])], target_guide=I()["mask_2_target"])
    S["c3_style_transfer_500"] = lambda: dict(examples=[ex(I()["multi4"])], target_guide=I()["tom"])
    # C4: inpaint + tiling, 1024^2 synthetic example, 2048^2 output (the session CatmullRom-resizes example and mask to the
    # inpaint dims, session.rs:346-379); and the non-inpaint reading "tiling 2048^2 from a 1024^2 example"
    S["c4_inpaint_tiling_2048"] = lambda: dict(examples=[ex(_synth(1024, 1024, 2))], inpaint=(_border_mask(2048, 2048), 0, (2048, 2048)),
                                               tiling=True)
    S["c4_tiling_2048_from_1024"] = lambda: dict(examples=[ex(_synth(1024, 1024, 2))], out=(2048, 2048), tiling=True)
    # C5: 8192^2 from the synthetic 1024^2 example
    S["c5_8192_from_1024"] = lambda: dict(examples=[ex(_synth(1024, 1024, 2))], out=(8192, 8192))
    # ---- the reference's own integration tests at the reference's sizes (lib/tests/diff.rs:163-252) --------------
    S["diff_single_example"] = lambda: dict(examples=[ex(I()["img1"])], seed=120, out=(100, 100))
    S["diff_multi_example"] = lambda: dict(examples=[ex(I()[f"multi{i}"]) for i in (1, 2, 3, 4)], resize=(100, 100),
                                           random_init=10, seed=211, out=(100, 100))
    S["diff_guided"] = lambda: dict(examples=[ex(I()["img2"], guide=I()["mask_2_example"])], target_guide=I()["mask_2_target"],
                                    out=(100, 100))
    S["diff_style_transfer"] = lambda: dict(examples=[ex(I()["multi4"])], target_guide=I()["tom"], out=(100, 100))
    S["diff_inpaint"] = lambda: dict(examples=[ex(I()["img3"], method="image", mask=I()["mask_3_inpaint"])],
                                     inpaint=(I()["mask_3_inpaint"], 0, (100, 100)))
    S["diff_inpaint_channel"] = lambda: dict(examples=[ex(I()["bricks"])], inpaint=(("channel", "A"), 0, (400, 400)))
    S["diff_tiling"] = lambda: dict(examples=[ex(I()["img1"])], inpaint=(I()["mask_1_tile"], 0, (100, 100)), tiling=True)
    S["diff_sample_masks"] = lambda: dict(examples=[ex(I()["img4"], method="image", mask=I()["mask_4_sample"])], seed=211,
                                          out=(100, 100))
    S["diff_sample_masks_ignore"] = lambda: dict(examples=[ex(I()["img4"], method="ignore"), ex(I()["img5"])], seed=211,
                                                 out=(200, 200))
    return S


SPECS = _specs()
DIFF_HASHES = {  # lib/tests/diff.rs:163-252
    "diff_single_example": "JKc2MqWo1iNWeJ856Ty6+a1M", "diff_multi_example": "JFCWyK1a4vJ1eWNTQkPOmdy2",
    "diff_guided": "JBQFEQoXm5CCiWZUfHHBhweK", "diff_style_transfer": "JEMRDSUzJ4uhpHMes1Onenz0",
    "diff_inpaint": "JNG1tl5SaIkqauco1NEmtikk", "diff_inpaint_channel": "JOVF4dThzPKa2suWLo1OWrKk",
    "diff_tiling": "JFSVUUmMaMzhWSttmlwojR1q", "diff_sample_masks": "JLO1hQBEpakECqIXDiCkqBME",
    "diff_sample_masks_ignore": "JGgWBEwJiqCaKpEiAonGkQRE",
}


def to_oracle(spec):
    """Spec -> tests.oracle_session.OracleSession (the checker)."""
    from oracle import ts_oracle as O
    from tests.oracle_session import OracleSession
    meth = {"all": O.METHOD_ALL, "ignore": O.METHOD_IGNORE, "image": O.METHOD_IMAGE}
    o = OracleSession()
    inp = spec.get("inpaint")
    for i, e in enumerate(spec["examples"]):
        if inp is not None and inp[1] == i:
            msk = inp[0]
            if isinstance(msk, tuple):
                o.inpaint_example_channel(msk[1], e["img"], inp[2])
            else:
                o.inpaint_example(msk, e["img"], inp[2], method=meth[e["method"]], sample_mask=e["mask"])
        else:
            o.add_example(e["img"], guide=e["guide"], method=meth[e["method"]], mask=e["mask"])
    o.target_guide = spec.get("target_guide")
    o.out_size = spec.get("out", (500, 500))
    o.resize = spec.get("resize")
    o.random_init_count = spec.get("random_init")
    o.p["seed"] = spec.get("seed", 0)
    o.p["tiling"] = spec.get("tiling", False)
    return o


def to_gpu(spec):
    """Spec -> texture_synthesis_b200.SessionBuilder (the product path, through the C ABI)."""
    import texture_synthesis_b200 as ts
    b = ts.Session.builder()
    inp = spec.get("inpaint")
    for i, e in enumerate(spec["examples"]):
        x = ts.Example(e["img"])
        if e["guide"] is not None:
            x.with_guide(e["guide"])
        if e["method"] == "ignore":
            x.set_sample_method(ts.SampleMethod.Ignore())
        elif e["method"] == "image":
            x.set_sample_method(e["mask"])
        if inp is not None and inp[1] == i:
            msk = inp[0]
            if isinstance(msk, tuple):
                b.inpaint_example_channel(msk[1], x, ts.Dims(*inp[2]))
            else:
                b.inpaint_example(msk, x, ts.Dims(*inp[2]))
        else:
            b.add_example(x)
    if spec.get("target_guide") is not None:
        b.load_target_guide(spec["target_guide"])
    if "out" in spec:
        b.output_size(ts.Dims(*spec["out"]))
    if spec.get("resize"):
        b.resize_input(ts.Dims(*spec["resize"]))
    if spec.get("random_init") is not None:
        b.random_init(spec["random_init"])
    b.seed(spec.get("seed", 0)).tiling_mode(spec.get("tiling", False)).max_thread_count(1)
    return b


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def digest(color, coord, ids, order, score, id_maps=None):
    """SHA-256 of every output the reference hands back: colour map, [x,y,map] coordinate transform, (patch id, map id),
    the resolution order and the first-resolution scores (bit patterns)."""
    d = dict(color=_sha(color), coord=_sha(coord), id=_sha(ids), order=_sha(np.asarray(order, np.uint32)),
             score=_sha(np.asarray(score, np.float32).view(np.uint32)), n_resolved=int(len(order)))
    if id_maps is not None:
        d["patch_id_png"], d["map_id_png"] = _sha(id_maps[0]), _sha(id_maps[1])
    return d


def digest_of_oracle(g, spec):
    flat, score = g.resolved()
    return digest(g.color(), g.coord(), g.ids(), flat, score, g.id_maps() if spec.get("id_maps") else None)


def digest_of_gpu(generated, spec):
    g = generated.inner
    flat, score = g.resolved()
    return digest(generated.into_image(), generated.get_coordinate_transform().buffer, g.ids(), flat, score,
                  g.id_maps() if spec.get("id_maps") else None)


def load_digests():
    if not os.path.exists(DIGESTS):
        return {}
    with open(DIGESTS) as f:
        return json.load(f)
