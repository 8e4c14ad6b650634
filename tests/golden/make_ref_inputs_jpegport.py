"""Second decode of the reference's JPEG inputs, with the restated jpeg-decoder 0.1.22 pixel pipeline (oracle/jpeg_port.py)
instead of Pillow's libjpeg, stored as int8 differences against tests/golden/ref_imgs_full.npz (they are 0 for > 99 % of the
samples and +-1..2 elsewhere, so the fixture is a few KiB).

    python tests/golden/make_ref_inputs_jpegport.py      (build container only: reads /root/reference/imgs)

Only the unsubsampled JPEGs are restated (everything the nine diff.rs configurations read except tom.jpg, 4:2:0, which stays a
Pillow decode).  Used by tests/test_oracle_pin.py through tests/fullsize_cases.py::set_decoder("jpegport")."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import jpeg_port  # noqa: E402

REF = "/root/reference/imgs"
FILES = {
    "img1": "1.jpg", "img2": "2.jpg", "img3": "3.jpg",
    "multi1": "multiexample/1.jpg", "multi2": "multiexample/2.jpg", "multi3": "multiexample/3.jpg", "multi4": "multiexample/4.jpg",
    "mask_1_tile": "masks/1_tile.jpg", "mask_2_example": "masks/2_example.jpg", "mask_2_target": "masks/2_target.jpg",
    "mask_3_inpaint": "masks/3_inpaint.jpg",
}


def main():
    base = np.load(os.path.join(HERE, "ref_imgs_full.npz"))
    out = {}
    for key, rel in FILES.items():
        a = jpeg_port.decode(os.path.join(REF, rel))
        b = base[key][..., :3]
        assert a.shape == b.shape, (key, a.shape, b.shape)
        d = a.astype(np.int16) - b.astype(np.int16)
        assert np.abs(d).max() <= 4, (key, np.abs(d).max())
        out[key] = d.astype(np.int8)
        print(f"{key}: {a.shape}, {(d != 0).mean() * 100:.3f} % of the samples differ from libjpeg's, max |d| = {np.abs(d).max()}")
    path = os.path.join(HERE, "ref_imgs_jpegport.npz")
    np.savez_compressed(path, **out)
    print(os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
