"""Restatement of the perceptual hash the reference's own integration tests pin their outputs with
(lib/tests/diff.rs:132-161: `ImageHash::hash(&MyPrecious(img), 8, HashType::DoubleGradient)` from the git-pinned fork
EmbarkStudios/img_hash@c40da78 of img_hash 2.1.0, Cargo.toml:7-11; the crate's source is not on disk).

TEST INFRASTRUCTURE ONLY (tests/ and the oracle pin report); nothing in the product imports this.

Algorithm (img_hash 2.x `double_gradient_hash`, with the reference's `HashImage` impl for `RgbaImage`, diff.rs:101-130):
  1. grayscale: image 0.23.12 `imageops::grayscale` -> luma = 0.2126 r + 0.7152 g + 0.0722 b in f32, truncating cast;
  2. resize to (hash_size + 1) x (hash_size + 1) = 9 x 9 with `FilterType::Nearest` (diff.rs:30): image 0.23.12's sampler with
     the box kernel and support 0 picks source pixel floor((o + 0.5) * n / 9);
  3. bits: for every row, `px[c-1] < px[c]` for c = 1..8 (9 x 8 = 72 bits), then for every column, `px[r-1][c] < px[r][c]`
     for r = 1..7 (9 x 7 = 63 bits) -- 135 bits;
  4. `to_base64`: one header byte (0x24 in every constant of diff.rs) + the bit vector packed MSB first, base64 without padding
     (18 bytes -> the 24-character constants).
The layout of step 3 is not documented anywhere reachable offline; it was identified from the reference's nine constants
themselves: with it, the oracle's outputs reproduce three of the nine hashes bit for bit, and every gradient between two
locked (never re-synthesised) pixels of the three inpainting configurations agrees with the constants (219 of 219 pairs).
"""
import base64

import numpy as np

HEADER = 0x24


def grayscale(rgba):
    r = rgba[..., 0].astype(np.float32)
    g = rgba[..., 1].astype(np.float32)
    b = rgba[..., 2].astype(np.float32)
    return ((np.float32(0.2126) * r + np.float32(0.7152) * g) + np.float32(0.0722) * b).astype(np.uint8)


def nearest_index(n, out):
    ratio = np.float32(n) / np.float32(out)
    return [min(int(np.floor((np.float32(o) + np.float32(0.5)) * ratio)), n - 1) for o in range(out)]


def sample_grid(rgba, hash_size=8):
    h, w = rgba.shape[:2]
    return nearest_index(h, hash_size + 1), nearest_index(w, hash_size + 1)


def bits(rgba, hash_size=8):
    ys, xs = sample_grid(rgba, hash_size)
    g = grayscale(rgba)[np.ix_(ys, xs)]
    n = hash_size + 1
    out = [int(g[r, c - 1] < g[r, c]) for r in range(n) for c in range(1, n)]
    out += [int(g[r - 1, c] < g[r, c]) for c in range(n) for r in range(1, hash_size)]
    return np.array(out, np.uint8)


def to_base64(b):
    packed = np.packbits(np.concatenate([b, np.zeros((-len(b)) % 8, np.uint8)]))
    return base64.b64encode(bytes([HEADER]) + packed.tobytes()).decode().rstrip("=")


def from_base64(s):
    raw = base64.b64decode(s + "=" * ((-len(s)) % 4))
    assert raw[0] == HEADER
    return np.unpackbits(np.frombuffer(raw[1:], np.uint8))


def hash_image(rgba):
    return to_base64(bits(rgba))


def distance(rgba, expected_b64):
    """Hamming distance between the hash of `rgba` and a constant of diff.rs (over the 135 hash bits)."""
    b = bits(rgba)
    return int((b != from_base64(expected_b64)[: len(b)]).sum())
