"""Decodes the reference's own test images ONCE (here, in the build container, with Pillow) into a small fixture.

    python tests/golden/make_ref_inputs.py

The GPU box has no /root/reference, and JPEG decoders differ by +-1 LSB (jpeg-decoder 0.1.22 vs libjpeg), so the
decoded bytes are frozen: the oracle and the CUDA path consume exactly the same arrays.  To keep the fixture
small every image is downscaled to at most 128 px on its longer side (Pillow LANCZOS); masks keep hard 0/255
values by thresholding after the resize.  Used by tests/test_gpu_reference_cases.py, which mirrors the nine
configurations of lib/tests/diff.rs:163-252.
"""
import os

import numpy as np
from PIL import Image

REF = "/root/reference/imgs"
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = {
    "img1": "1.jpg", "img2": "2.jpg", "img3": "3.jpg", "img4": "4.png", "img5": "5.png", "bricks": "bricks.png", "tom": "tom.jpg",
    "multi1": "multiexample/1.jpg", "multi2": "multiexample/2.jpg", "multi3": "multiexample/3.jpg", "multi4": "multiexample/4.jpg",
    "mask_1_tile": "masks/1_tile.jpg", "mask_2_example": "masks/2_example.jpg", "mask_2_target": "masks/2_target.jpg",
    "mask_3_inpaint": "masks/3_inpaint.jpg", "mask_4_sample": "masks/4_sample_mask.png",
}


def main():
    out = {}
    for key, rel in FILES.items():
        im = Image.open(os.path.join(REF, rel)).convert("RGBA")
        scale = 128.0 / max(im.size)
        if scale < 1.0:
            im = im.resize((max(1, round(im.size[0] * scale)), max(1, round(im.size[1] * scale))), Image.LANCZOS)
        a = np.asarray(im).copy()
        if key.startswith("mask_1") or key.startswith("mask_3") or key.startswith("mask_4"):
            v = np.where(a[..., 0] >= 128, 255, 0).astype(np.uint8)   # hard masks (ms.rs:272 needs R == 255, ms.rs:1546 R != 0)
            a[..., 0] = a[..., 1] = a[..., 2] = v
            a[..., 3] = 255
        out[key] = a
        print(key, a.shape)
    path = os.path.join(HERE, "ref_imgs.npz")
    np.savez_compressed(path, **out)
    print(os.path.getsize(path) // 1024, "KiB")
    # full-size decodes (no rescale, masks untouched: the reference thresholds the decoded bytes itself, ms.rs:272,1546),
    # consumed by tests/fullsize_cases.py for the BASELINE configs and lib/tests/diff.rs at the reference's real sizes.
    # RGB only where alpha is 255 everywhere (smaller file).
    full = {}
    for key, rel in FILES.items():
        a = np.asarray(Image.open(os.path.join(REF, rel)).convert("RGBA")).copy()
        full[key] = a[..., :3].copy() if (a[..., 3] == 255).all() else a
    path = os.path.join(HERE, "ref_imgs_full.npz")
    np.savez_compressed(path, **full)
    print(os.path.getsize(path) // 1024, "KiB (full size)")


if __name__ == "__main__":
    main()
