"""Self-contained checks of the two crate restatements behind the oracle pin (tests/test_oracle_pin.py): the R*-tree port answers
k-NN queries correctly and keeps rstar's node sizes; the JPEG port decodes baseline and progressive files produced here by Pillow
to within the +-3 that separates jpeg-decoder's IDCT / colour conversion from libjpeg's.  CPU only, no reference files."""
import io
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rstar_port_is_a_correct_rstar_tree(tmp_path):
    exe = str(tmp_path / "rstar_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "cpp", "rstar_check.cpp"),
                    "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "bad queries 0" in r.stdout


@pytest.mark.parametrize("progressive", [False, True])
@pytest.mark.parametrize("quality", [75, 95])
def test_jpeg_port_decodes_what_libjpeg_decodes(progressive, quality):
    Image = pytest.importorskip("PIL.Image")
    from oracle import jpeg_port
    rng = np.random.default_rng(7)
    yy, xx = np.mgrid[0:83, 0:101]
    img = np.stack([(xx * 2 + yy) % 256, (yy * 3) % 256, (xx + yy * 2) % 256], axis=-1).astype(np.uint8)
    img[20:50, 30:70] = rng.integers(0, 256, (30, 40, 3), dtype=np.uint8)     # some high-frequency content
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, "JPEG", quality=quality, subsampling=0, progressive=progressive, optimize=progressive)
    data = buf.getvalue()
    got = jpeg_port.decode(data)
    want = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    assert got.shape == want.shape == img.shape
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= 3, d.max()               # same coefficients; only IDCT / colour rounding differ
    assert (d != 0).mean() < 0.15
