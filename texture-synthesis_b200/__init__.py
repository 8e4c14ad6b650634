"""texture-synthesis_b200: the per-pixel nearest-neighbour patch search of EmbarkStudios/texture-synthesis
as hand-written CUDA for B200 (sm_100a) behind a C ABI (include/tsb200.h).

`capi`    -- ctypes binding of libtsb200.so (the product; no CPU fallback)
`session` -- host-side mirror of the reference's Session::builder() API for this path
"""
from . import capi  # noqa: F401
from .session import (CoordinateTransform, Dims, Error, Example, GeneratedImage, InvalidRange, SampleMethod,  # noqa: F401
                      Session, SessionBuilder, load_image)
