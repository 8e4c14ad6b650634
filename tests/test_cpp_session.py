"""The C++ host-side mirror of the reference API (include/tsb200_session.hpp): builds with g++ against libtsb200.so;
parameter validation runs on CPU, the synthesis flows on the GPU and must match the Python mirror byte for byte."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "texture-synthesis_b200")


def _build(tmp_path):
    exe = str(tmp_path / "session_example")
    cmd = ["g++", "-std=c++17", "-O2", "-w", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "session_example.cpp"),
           "-o", exe, "-L" + LIBDIR, "-ltsb200", "-Wl,-rpath," + LIBDIR]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def test_cpp_mirror_builds_and_validates_like_the_reference(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe, "validate"], capture_output=True, text=True)
    assert out.returncode == 0 and "validate 4/4" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["single", "tiling", "style"])
def test_cpp_mirror_matches_python_mirror(tmp_path, mode):
    sys.path.insert(0, ROOT)
    import texture_synthesis_b200 as ts
    from texture_synthesis_b200.synth import synth_texture
    exe = _build(tmp_path)
    ex, tgt = synth_texture(64, 56, 21), synth_texture(80, 72, 22)
    (tmp_path / "ex.rgba").write_bytes(ex.tobytes())
    (tmp_path / "tgt.rgba").write_bytes(tgt.tobytes())
    args = [exe, mode, str(tmp_path / "ex.rgba"), "64", "56", str(tmp_path / "out.rgba"), "80", "72"]
    if mode == "style":
        args += [str(tmp_path / "tgt.rgba"), "80", "72"]
    out = subprocess.run(args, capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr
    got = np.frombuffer((tmp_path / "out.rgba").read_bytes(), np.uint8).reshape(72, 80, 4)
    b = ts.Session.builder().add_example(ex).seed(120).output_size(ts.Dims(80, 72)).max_thread_count(1)
    if mode == "tiling":
        b.tiling_mode(True)
    if mode == "style":
        b.load_target_guide(tgt)
    want = b.build().run(None).into_image()
    assert (got == want).all()
