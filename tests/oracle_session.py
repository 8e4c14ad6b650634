"""The reference's Session::build + Session::run pipeline (lib/src/session.rs:37-66, 336-448; lib.rs:564-603) driven
entirely by the CPU oracle: the checker for tests that go through the Session mirrors.  Accepts the same builder
calls as texture_synthesis_b200.session.SessionBuilder."""
import numpy as np

from oracle import ts_oracle as O


def _load(img, size):
    """utils::load_image (utils.rs:55-80) for a decoded RGBA array: CatmullRom resize if the size differs."""
    if size is not None and (img.shape[1], img.shape[0]) != tuple(size):
        return O.resize(img, size[0], size[1], O.F_CATMULLROM)
    return np.ascontiguousarray(img, np.uint8)


class OracleSession:
    def __init__(self):
        self.examples = []          # dicts: img, guide, method, mask
        self.target_guide = None
        self.inpaint = None         # (mask_img or ('channel', ch), example_index, (w, h))
        self.p = dict(k=50, m=50, cauchy=1.0, p=0.5, stages=5, seed=0, alpha=0.8, tiling=False)
        self.out_size = (500, 500)
        self.resize = None
        self.random_init_count = None

    def add_example(self, img, guide=None, method=O.METHOD_ALL, mask=None):
        self.examples.append(dict(img=img, guide=guide, method=method, mask=mask))
        return self

    def inpaint_example(self, mask, img, size, method=O.METHOD_ALL, sample_mask=None):
        self.inpaint = (mask, len(self.examples), size)
        return self.add_example(img, None, method, sample_mask)

    def inpaint_example_channel(self, channel, img, size):
        self.inpaint = (("channel", channel), len(self.examples), size)
        return self.add_example(img)

    def run(self):
        levels = self.p["stages"]
        if self.inpaint is not None:
            msrc, ex_index, dims = self.inpaint
            if isinstance(msrc, tuple):  # utils::apply_mask (utils.rs:82-99)
                base = _load(self.examples[ex_index]["img"], dims)
                ch = {"R": 0, "G": 1, "B": 2, "A": 3}[msrc[1]]
                mask_img = np.ascontiguousarray(np.stack([base[..., ch]] * 3 + [np.full(base.shape[:2], 255, np.uint8)], axis=-1))
            else:
                mask_img = _load(msrc, dims)
            color = _load(self.examples[ex_index]["img"], dims)
            out_size, in_size = dims, dims
        else:
            mask_img = color = None
            ex_index = 0
            out_size, in_size = self.out_size, self.resize
        target_pyr = None
        if self.target_guide is not None:
            tg = _load(self.target_guide, out_size)
            if not any(e["guide"] is not None for e in self.examples):
                tg = O.guide_map(tg, 2.0)
            target_pyr = O.pyramid_build(tg, levels)
        pyrs, gpyrs, methods, masks = [], [], [], []
        for e in self.examples:
            pyr = O.pyramid_build(_load(e["img"], in_size), levels)
            pyrs.append(pyr)
            if target_pyr is not None:
                if e["guide"] is not None:
                    gpyrs.append(O.pyramid_build(_load(e["guide"], in_size), levels))
                else:
                    gpyrs.append(O.pyramid_build(O.match_histograms(O.guide_map(pyr[-1], 2.0), target_pyr[-1]), levels))
            methods.append(e["method"])
            masks.append(_load(e["mask"], in_size) if e["method"] == O.METHOD_IMAGE else None)
        g = O.Generator(out_size[0], out_size[1], mask_img, color, ex_index)
        g.set_examples(pyrs, methods, masks)
        if target_pyr is not None:
            g.set_guides(target_pyr, gpyrs)
        if self.random_init_count is not None:
            g.random_init(self.random_init_count, self.p["seed"])
        g.resolve(O.make_params(k=self.p["k"], m=self.p["m"], cauchy=self.p["cauchy"], p=self.p["p"], stages=levels,
                                seed=self.p["seed"], alpha=self.p["alpha"], threads=1, tiling=self.p["tiling"]))
        return g
