"""pixo_b200 — B200 (sm_100a) drop-in for the data-parallel stages of pixo's JPEG/PNG encoders.

Python mirror of the reference's public API for this path:
  pixo_b200.jpeg  <->  pixo::jpeg   (encode, encode_into, JpegOptions, Subsampling)
  pixo_b200.png   <->  pixo::png    (filter::apply_filters*, FilterStrategy, PngOptions) and
                       pixo::compress::adler32
All arithmetic happens in libpixo_b200.so (hand-written CUDA behind a C ABI, include/pixo_b200.h).
"""
from ._lib import PixoError, SO_PATH, load  # noqa: F401
from .color import ColorType  # noqa: F401
from .context import Context, default_context  # noqa: F401
from . import jpeg, png  # noqa: F401

__all__ = ["PixoError", "ColorType", "Context", "default_context", "jpeg", "png", "load", "SO_PATH"]
