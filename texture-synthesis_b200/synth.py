"""Deterministic synthetic example textures for the benchmark configs.

synth_texture(w, h, seed): per colour channel three octaves of bilinear value noise (lattice
periods 64/16/4 px, amplitudes 96/48/24 around 128); lattice values are successive `next_u32()`
outputs of Pcg32::seed_from_u64(seed*16 + channel*4 + octave) mapped to [-0.5, 0.5); A = 255.
"""
import hashlib

import numpy as np

from .rng import Pcg32

PERIODS = (64, 16, 4)
AMPS = (96.0, 48.0, 24.0)


def _lattice(seed, ny, nx):
    rng = Pcg32.seed_from_u64(seed)
    v = np.array([rng.next_u32() for _ in range(ny * nx)], dtype=np.float64)
    return (v / 4294967296.0 - 0.5).reshape(ny, nx)


def synth_texture(w, h, seed):
    img = np.empty((h, w, 4), np.uint8)
    img[..., 3] = 255
    ys = np.arange(h, dtype=np.float64)
    xs = np.arange(w, dtype=np.float64)
    for ch in range(3):
        acc = np.full((h, w), 128.0)
        for o, (per, amp) in enumerate(zip(PERIODS, AMPS)):
            ny, nx = h // per + 2, w // per + 2
            lat = _lattice(seed * 16 + ch * 4 + o, ny, nx)
            fy, fx = ys / per, xs / per
            iy, ix = fy.astype(np.int64), fx.astype(np.int64)
            ty, tx = (fy - iy)[:, None], (fx - ix)[None, :]
            a = lat[iy][:, ix]
            b = lat[iy][:, ix + 1]
            c = lat[iy + 1][:, ix]
            d = lat[iy + 1][:, ix + 1]
            acc += amp * ((a * (1 - tx) + b * tx) * (1 - ty) + (c * (1 - tx) + d * tx) * ty)
        img[..., ch] = np.clip(np.floor(acc), 0, 255).astype(np.uint8)
    return img


def border_inpaint_mask(w, h, frac=0.17):
    """C4 mask: a border band of width floor(frac*w) is 0 (to synthesise), interior 255 (kept)."""
    m = np.zeros((h, w, 4), np.uint8)
    m[..., 3] = 255
    bx, by = int(frac * w), int(frac * h)
    m[by:h - by, bx:w - bx, :3] = 255
    return m


def sha256(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
