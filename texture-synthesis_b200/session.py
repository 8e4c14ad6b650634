"""Host-side mirror of the reference's public API for the hot path, on top of the C ABI.

The reference is Rust (`Session::builder() ... build() ... run()`, lib/src/session.rs:32-448); no Rust
toolchain exists in the build image, so this module restates that interface in Python with the same
names, argument meaning and error behaviour, and INTEGRATION.md shows the Rust `extern "C"` shim a
maintainer would add.  All pixel work goes through libtsb200.so (CUDA); nothing here falls back to a CPU
implementation of the synthesis.

    import texture_synthesis_b200 as ts
    sess = ts.Session.builder().add_example(img).seed(10).tiling_mode(True).build()
    generated = sess.run(None)
    generated.save("out.png")
"""
import struct

import numpy as np

from . import capi


class Error(Exception):
    """Mirror of texture_synthesis::Error (lib/src/errors.rs:38-57); `kind` names the variant."""

    def __init__(self, kind, msg):
        super().__init__(msg)
        self.kind = kind


class InvalidRange(Error):
    def __init__(self, name, lo, hi, value):
        super().__init__("InvalidRange", f"parameter '{name}' - value '{value}' is outside the range of {lo}-{hi}")
        self.name, self.min, self.max, self.value = name, lo, hi, value


class Dims:
    """lib/src/lib.rs:142-160"""

    def __init__(self, width, height):
        self.width, self.height = int(width), int(height)

    @staticmethod
    def square(size):
        return Dims(size, size)

    def __eq__(self, o):
        return isinstance(o, Dims) and (self.width, self.height) == (o.width, o.height)

    def __repr__(self):
        return f"Dims({self.width}x{self.height})"


def _load_rgba(src):
    """ImageSource (utils.rs:6-48): a path, encoded bytes or an already decoded array."""
    if isinstance(src, np.ndarray):
        a = src
        if a.ndim == 2:
            a = np.stack([a, a, a, np.full_like(a, 255)], axis=-1)
        elif a.shape[2] == 3:
            a = np.concatenate([a, np.full(a.shape[:2] + (1,), 255, a.dtype)], axis=-1)
        return np.ascontiguousarray(a, np.uint8)
    try:
        from PIL import Image
        import io
        im = Image.open(io.BytesIO(src)) if isinstance(src, (bytes, bytearray)) else Image.open(src)
        return np.ascontiguousarray(np.asarray(im.convert("RGBA")), np.uint8)
    except Exception as e:  # image::ImageError -> Error::Image
        raise Error("Image", str(e))


def load_image(src, resize=None):
    """utils.rs:55-80: decode to RGBA8 and CatmullRom-resize to `resize` if the size differs."""
    img = _load_rgba(src)
    if resize is not None and (img.shape[1], img.shape[0]) != (resize.width, resize.height):
        img = capi.resize(img, resize.width, resize.height, capi.FILTER_CATMULLROM)
    return img


class SampleMethod:
    """lib.rs:458-484"""
    ALL, IGNORE, IMAGE = capi.SAMPLE_ALL, capi.SAMPLE_IGNORE, capi.SAMPLE_IMAGE

    def __init__(self, kind, img=None):
        self.kind, self.img = kind, img

    @staticmethod
    def All():
        return SampleMethod(SampleMethod.ALL)

    @staticmethod
    def Ignore():
        return SampleMethod(SampleMethod.IGNORE)

    @staticmethod
    def Image(src):
        return SampleMethod(SampleMethod.IMAGE, src)


class Example:
    """lib.rs:487-623 (Example / ExampleBuilder)."""

    def __init__(self, img):
        self.img, self.guide, self.sample_method = img, None, SampleMethod.All()

    @staticmethod
    def builder(img):
        return Example(img)

    def with_guide(self, guide):
        self.guide = guide
        return self

    def set_sample_method(self, method):
        self.sample_method = method if isinstance(method, SampleMethod) else SampleMethod.Image(method)
        return self


class _Params:
    """lib.rs:327-359 defaults"""

    def __init__(self):
        self.tiling_mode = False
        self.nearest_neighbors = 50
        self.random_sample_locations = 50
        self.cauchy_dispersion = 1.0
        self.backtrack_percent = 0.5
        self.backtrack_stages = 5
        self.resize_input = None
        self.output_size = Dims.square(500)
        self.guide_alpha = 0.8
        self.random_resolve = None
        self.max_thread_count = None
        self.seed = 0


class SessionBuilder:
    """lib/src/session.rs:72-524"""

    def __init__(self):
        self.examples = []
        self.target_guide = None
        self.inpaint_mask = None  # (mask source or channel, example index, dims)
        self.params = _Params()

    def add_example(self, example):
        self.examples.append(example if isinstance(example, Example) else Example(example))
        return self

    def add_examples(self, examples):
        for e in examples:
            self.add_example(e)
        return self

    def inpaint_example(self, inpaint_mask, example, size):
        self.inpaint_mask = (("img", inpaint_mask), len(self.examples), size)
        return self.add_example(example)

    def inpaint_example_channel(self, mask_channel, example, size):
        """mask_channel: one of 'R','G','B','A' (utils.rs:50-53, 82-99)."""
        self.inpaint_mask = (("channel", mask_channel), len(self.examples), size)
        return self.add_example(example)

    def load_target_guide(self, guide):
        self.target_guide = guide
        return self

    def resize_input(self, dims):
        self.params.resize_input = dims
        return self

    def seed(self, value):
        self.params.seed = int(value)
        return self

    def tiling_mode(self, is_tiling):
        self.params.tiling_mode = bool(is_tiling)
        return self

    def nearest_neighbors(self, count):
        self.params.nearest_neighbors = int(count)
        return self

    def random_sample_locations(self, count):
        self.params.random_sample_locations = int(count)
        return self

    def random_init(self, count):
        self.params.random_resolve = int(count)
        return self

    def cauchy_dispersion(self, value):
        self.params.cauchy_dispersion = float(value)
        return self

    def guide_alpha(self, value):
        self.params.guide_alpha = float(value)
        return self

    def backtrack_percent(self, value):
        self.params.backtrack_percent = float(value)
        return self

    def backtrack_stages(self, stages):
        self.params.backtrack_stages = int(stages)
        return self

    def output_size(self, dims):
        self.params.output_size = dims
        return self

    def max_thread_count(self, count):
        self.params.max_thread_count = int(count)
        return self

    # session.rs:450-499
    def _check_parameters_validity(self):
        p = self.params
        if not (0.0 <= p.cauchy_dispersion <= 1.0):
            raise InvalidRange("cauchy-dispersion", 0.0, 1.0, p.cauchy_dispersion)
        if not (0.0 <= p.backtrack_percent <= 1.0):
            raise InvalidRange("backtrack-percent", 0.0, 1.0, p.backtrack_percent)
        if not (0.0 <= p.guide_alpha <= 1.0):
            raise InvalidRange("guide-alpha", 0.0, 1.0, p.guide_alpha)
        if p.max_thread_count is not None and p.max_thread_count == 0:
            raise InvalidRange("max-thread-count", 1.0, 1024.0, 0.0)
        if p.random_sample_locations == 0:
            raise InvalidRange("m-rand", 1.0, 1024.0, 0.0)

    # session.rs:501-524
    def _check_images_validity(self):
        usable = [e for e in self.examples if e.sample_method.kind != SampleMethod.IGNORE]
        if not usable:
            raise Error("NoExamples", "at least 1 example that is not ignored is required")
        n_guides = sum(1 for e in self.examples if e.guide is not None)
        if n_guides != 0 and n_guides != len(self.examples):
            raise Error("ExampleGuideMismatch", f"{len(self.examples)} examples but {n_guides} guides")

    def build(self):
        """session.rs:336-448"""
        self._check_parameters_validity()
        self._check_images_validity()
        p = self.params
        levels = p.backtrack_stages
        inpaint = None
        if self.inpaint_mask is not None:
            (kind, src), ex_index, dims = self.inpaint_mask
            if kind == "img":
                mask_img = load_image(src, dims)
            else:  # utils::apply_mask (utils.rs:82-99)
                base = load_image(self.examples[ex_index].img, dims)
                ch = {"R": 0, "G": 1, "B": 2, "A": 3}[src]
                mask_img = np.ascontiguousarray(np.stack([base[..., ch]] * 3 + [np.full(base.shape[:2], 255, np.uint8)], axis=-1))
            color = load_image(self.examples[ex_index].img, dims)
            inpaint = (mask_img, color, ex_index)
            out_size, in_size = dims, dims
        else:
            out_size, in_size = p.output_size, p.resize_input

        target_pyr = None
        if self.target_guide is not None:
            tg = load_image(self.target_guide, out_size)
            if not any(e.guide is not None for e in self.examples):
                tg = transform_to_guide_map(tg, 2.0)
            target_pyr = capi.pyramid_build(tg, levels)

        pyramids, guide_pyrs, methods, masks = [], [], [], []
        for e in self.examples:
            img = load_image(e.img, in_size)
            pyr = capi.pyramid_build(img, levels)  # ImagePyramid::new (lib.rs:570)
            pyramids.append(pyr)
            if target_pyr is not None:
                if e.guide is not None:
                    guide_pyrs.append(capi.pyramid_build(load_image(e.guide, in_size), levels))
                else:  # lib.rs:577-581
                    gm = transform_to_guide_map(pyr[-1].copy(), 2.0)
                    gm = match_histograms(gm, target_pyr[-1])
                    guide_pyrs.append(capi.pyramid_build(gm, levels))
            methods.append(e.sample_method.kind)
            masks.append(load_image(e.sample_method.img, in_size) if e.sample_method.kind == SampleMethod.IMAGE else None)

        gen = capi.Generator(out_size.width, out_size.height,
                             inpaint[0] if inpaint else None, inpaint[1] if inpaint else None,
                             inpaint[2] if inpaint else 0)
        guides = (target_pyr, guide_pyrs) if target_pyr is not None else None
        return Session(pyramids, guides, methods, masks, gen, p)


class Session:
    """lib/src/session.rs:22-66"""

    def __init__(self, examples, guides, methods, masks, generator, params):
        self.examples, self.guides, self.methods, self.masks = examples, guides, methods, masks
        self.generator, self.params = generator, params

    @staticmethod
    def builder():
        return SessionBuilder()

    def generator_params(self):
        p = self.params
        return capi.make_params(k=p.nearest_neighbors, m=p.random_sample_locations, cauchy=p.cauchy_dispersion,
                                p=p.backtrack_percent, stages=p.backtrack_stages, seed=p.seed, alpha=p.guide_alpha,
                                threads=p.max_thread_count or 1, tiling=p.tiling_mode)

    def run(self, progress=None):
        """progress(image, (total_current, total_total), (stage_current, stage_total)) mirrors GeneratorProgress."""
        p = self.params
        if p.random_resolve is not None:  # session.rs:42-52
            self.generator.random_init(p.random_resolve, [pyr[-1] for pyr in self.examples], p.seed)
        self.generator.resolve(self.generator_params(), self.examples, self.methods, self.masks, self.guides, progress)
        dims = [Dims(pyr.shape[2], pyr.shape[1]) for pyr in self.examples]
        return GeneratedImage(self.generator, dims)


class CoordinateTransform:
    """lib/src/lib.rs:162-325: [x, y, map] u32 triplets + output size + original map sizes."""
    MAGIC = 0x1234_0001

    def __init__(self, buffer, output_size, original_maps):
        self.buffer, self.output_size, self.original_maps = buffer, output_size, original_maps

    def write(self, fp):
        header = [self.MAGIC, self.output_size.width, self.output_size.height, len(self.original_maps)]
        for d in self.original_maps:
            header += [d.width, d.height]
        fp.write(struct.pack(f"={len(header)}I", *header))
        fp.write(np.ascontiguousarray(self.buffer, np.uint32).tobytes())

    @staticmethod
    def read(fp):
        def u32():
            b = fp.read(4)
            if len(b) != 4:
                raise Error("Io", "unexpected end of coordinate transform")
            return struct.unpack("=I", b)[0]
        if u32() != CoordinateTransform.MAGIC:
            raise Error("Io", "invalid magic")
        w, h, n = u32(), u32(), u32()
        maps = [Dims(u32(), u32()) for _ in range(n)]
        data = fp.read(w * h * 3 * 4)
        if len(data) != w * h * 3 * 4:
            raise Error("Io", "unexpected end of coordinate transform")
        return CoordinateTransform(np.frombuffer(data, np.uint32).reshape(h, w, 3).copy(), Dims(w, h), maps)

    def apply(self, sources):
        """lib.rs:178-210: re-synthesise from new source images (resized to the recorded sizes)."""
        if len(sources) != len(self.original_maps):
            raise Error("MapsCountMismatch", f"{len(sources)} inputs for {len(self.original_maps)} maps")
        imgs = [load_image(s, d) for s, d in zip(sources, self.original_maps)]
        out = np.zeros((self.output_size.height, self.output_size.width, 4), np.uint8)
        b = self.buffer.reshape(self.output_size.height, self.output_size.width, 3)
        for m, img in enumerate(imgs):
            sel = b[..., 2] == m
            out[sel] = img[b[..., 1][sel], b[..., 0][sel]]
        return out


class GeneratedImage:
    """lib/src/lib.rs:378-455"""

    def __init__(self, generator, input_dims):
        self.inner = generator
        self._input_dims = input_dims

    def into_image(self):
        return self.inner.color()

    def as_array(self):
        return self.inner.color()

    def save(self, path):
        from PIL import Image
        img = self.inner.color()
        im = Image.fromarray(img, "RGBA")
        if str(path).lower().endswith((".jpg", ".jpeg")):
            im = im.convert("RGB")
        im.save(path)

    def save_debug(self, dir_):
        """lib.rs:407-419: uncertainty.png, patch_id.png, map_id.png"""
        import os
        from PIL import Image
        os.makedirs(dir_, exist_ok=True)
        Image.fromarray(self.inner.uncertainty_map(), "RGBA").save(os.path.join(dir_, "uncertainty.png"))
        patch, maps = self.inner.id_maps()
        Image.fromarray(patch, "RGBA").save(os.path.join(dir_, "patch_id.png"))
        Image.fromarray(maps, "RGBA").save(os.path.join(dir_, "map_id.png"))

    def get_coordinate_transform(self):
        return CoordinateTransform(self.inner.coord(), Dims(self.inner.W, self.inner.H), list(self._input_dims))


# ---- guide preprocessing (utils.rs:101-183) runs on the GPU (tsb_guide_map / tsb_match_histograms) ----------
def transform_to_guide_map(img, blur_sigma):
    """utils.rs:101-116: blur(sigma) -> grayscale -> RGBA (the resize in the reference is a discarded no-op, q10)."""
    return capi.guide_map(img, blur_sigma)


def match_histograms(source, target):
    """utils.rs:135-183"""
    return capi.match_histograms(source, target)
