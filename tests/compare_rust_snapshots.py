"""Compares snapshots dumped by the REAL reference crate (rust/patches/lib/tests/dump_snapshots.rs, run by somebody with a Rust
toolchain) with the CUDA path and with the CPU oracle, on exactly the decoded inputs the Rust `image` crate produced:

    python tests/compare_rust_snapshots.py <dir with one sub-directory per diff.rs configuration> [--oracle-only]

Reports, per configuration, the mismatched-pixel fraction of the output image and of the coordinate transform (the end-to-end
criterion of BASELINE.json's north_star), three ways:
  * oracle with the restated rstar neighbour order (ORC_KNN=rstar) vs the crate -- expected 0 everywhere: this is the
    configuration that reproduces all nine hash constants of lib/tests/diff.rs (tests/test_oracle_pin.py);
  * oracle with the canonical order vs the crate, and the CUDA path (bit-identical to it) vs the crate -- expected to differ
    where equidistant neighbours straddle the k-cut (DESIGN.md section 2 has the table).
Not a pytest file: the snapshots cannot be produced in the build image."""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import fullsize_cases as F  # noqa: E402


def rgba(d, name):
    p = os.path.join(d, name + ".rgba")
    if not os.path.exists(p):
        return None
    w, h = map(int, open(os.path.join(d, name + ".dims")).read().split())
    return np.fromfile(p, np.uint8).reshape(h, w, 4)


def transform(path):
    """CoordinateTransform::write layout (lib.rs:212-249): magic, w, h, n_maps, n_maps x (w, h), then w*h*3 u32."""
    raw = open(path, "rb").read()
    magic, w, h, n = struct.unpack_from("=4I", raw, 0)
    assert magic == 0x12340001
    return np.frombuffer(raw, np.uint32, w * h * 3, 16 + n * 8).reshape(h, w, 3)


def spec_for(name, d):
    """The spec of tests/fullsize_cases.py for this configuration with the Rust-decoded inputs substituted."""
    spec = F.SPECS["diff_" + name]()
    for i, e in enumerate(spec["examples"]):
        e["img"] = rgba(d, f"input_{i}")
        g = rgba(d, f"guide_{i}")
        if g is not None:
            e["guide"] = g
        if e["mask"] is not None:
            e["mask"] = rgba(d, "mask")
    if spec.get("target_guide") is not None:
        spec["target_guide"] = rgba(d, "target_guide")
    if spec.get("inpaint") is not None and not isinstance(spec["inpaint"][0], tuple):
        spec["inpaint"] = (rgba(d, "mask"),) + tuple(spec["inpaint"][1:])
    return spec


def main():
    root = sys.argv[1]
    oracle_only = "--oracle-only" in sys.argv
    for name in sorted(os.listdir(root)):
        d = os.path.join(root, name)
        if not os.path.isdir(d) or "diff_" + name not in F.SPECS:
            continue
        want_img, want_tx = rgba(d, "output"), transform(os.path.join(d, "transform.bin"))
        spec = spec_for(name, d)
        os.environ["ORC_KNN"] = "rstar"
        r = F.to_oracle(spec).run()
        os.environ.pop("ORC_KNN")
        o = F.to_oracle(spec).run()
        line = (f"{name:22s} oracle (rstar order) vs crate: colour mismatch {np.mean((r.color() != want_img).any(axis=2)):.4f}, coord mismatch "
                f"{np.mean((r.coord() != want_tx).any(axis=2)):.4f} | oracle (canonical) vs crate: colour {np.mean((o.color() != want_img).any(axis=2)):.4f}, "
                f"coord {np.mean((o.coord() != want_tx).any(axis=2)):.4f}")
        if not oracle_only:
            g = F.to_gpu(spec).build().run(None)
            line += f" | CUDA vs crate: colour {np.mean((g.into_image() != want_img).any(axis=2)):.4f}, coord {np.mean((g.get_coordinate_transform().buffer != want_tx).any(axis=2)):.4f}"
        print(line, flush=True)


if __name__ == "__main__":
    main()
