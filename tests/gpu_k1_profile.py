"""K1 at C5 scale (not a pytest file): the Gaussian pyramid of an 8192x8192 target guide (img_pyramid.rs:20-37, 5 levels,
reduction factors 16, 8, 4, 2; the /16 level has 97-tap windows), timed through the C ABI and, under ncu, per kernel:
    python tests/gpu_k1_profile.py [size]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from texture_synthesis_b200 import capi
from texture_synthesis_b200.synth import synth_texture
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
img = np.tile(synth_texture(1024, 1024, 3), (n // 1024, n // 1024, 1))
for rep in range(2):
    t0 = time.time(); pyr = capi.pyramid_build(img, 5); dt = time.time() - t0
    print(f"pyramid {n}^2 x 5 levels: {dt * 1e3:.1f} ms wall (H2D {img.nbytes >> 20} MiB + D2H {pyr.nbytes >> 20} MiB inside)", flush=True)
