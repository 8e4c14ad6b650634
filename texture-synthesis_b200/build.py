"""Builds libtsb200.so (the C-ABI CUDA library) in-tree for sm_100a."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "tsb200.cu")
DEPS = [SRC] + [os.path.join(HERE, "csrc", f) for f in ("tsb_device.cuh", "tsb_stream.cuh", "tsb_rng.cuh")] + \
       [os.path.join(HERE, "..", "include", "tsb200.h")]
OUT = os.path.join(HERE, "libtsb200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--fmad=false",                      # the reference's f32/f64 arithmetic is unfused (ms.rs:1259-1280)
    "-Xcompiler", "-fPIC,-O2,-fno-fast-math,-ffp-contract=off", "-shared", "-Xptxas", "-v", "-ldl",
]


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    tmp = OUT + ".tmp%d" % os.getpid()  # built next to the target and renamed: a snapshot never sees a half-written library
    cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed")
    os.replace(tmp, OUT)
    with open(os.path.join(HERE, "csrc", "ptxas_info.txt"), "w") as f:
        f.write(r.stderr)
    return OUT


if __name__ == "__main__":
    build(force=True, verbose=True)
