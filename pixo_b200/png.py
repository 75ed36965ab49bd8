"""Mirror of pixo::png's filter stage and pixo::compress::adler32 for the accelerated path.

  apply_filters / apply_filters_with_row_bytes   src/png/filter.rs:52-206
  FilterStrategy                                 src/png/mod.rs:345-364
  adler32                                        src/compress/adler32.rs:11
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import enum

import numpy as np

from . import _lib
from .color import ColorType
from .context import Context, default_context


class FilterStrategy(enum.IntEnum):
    NoFilter = 0  # FilterStrategy::None
    Sub = 1
    Up = 2
    Average = 3
    Paeth = 4
    MinSum = 5
    Adaptive = 6
    AdaptiveFast = 7
    Bigrams = 8


OPTIMIZE_ALPHA = 0x100  # PIXO_B200_PNG_OPTIMIZE_ALPHA


@dataclasses.dataclass
class PngOptions:
    """The fields of pixo::png::PngOptions (src/png/mod.rs:41-118) the filter stage reads."""
    width: int = 0
    height: int = 0
    color_type: ColorType = ColorType.Rgba
    filter_strategy: FilterStrategy = FilterStrategy.Adaptive
    optimize_alpha: bool = False   # applied on the fly (Rgba / GrayAlpha), src/png/mod.rs:633-671


def _as_u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    return np.frombuffer(data, dtype=np.uint8)


def apply_filters_with_row_bytes(data, width, height, row_bytes, bytes_per_pixel, options: PngOptions,
                                 with_adler=False, ctx: Context | None = None):
    """filter::apply_filters_with_row_bytes (src/png/filter.rs:64): filtered rows, each
    prefixed by its filter-type byte; optionally also the Adler-32 of that stream."""
    ctx = ctx or default_context()
    d = _as_u8(data)
    if d.size != int(row_bytes) * int(height):
        raise _lib.PixoError(_lib.ERR_INVALID_DATA_LENGTH,
                             f"Invalid data length: expected {int(row_bytes) * int(height)} bytes, got {d.size}")
    out = np.empty(int(height) * (int(row_bytes) + 1), np.uint8)
    ad = C.c_uint32()
    rc = _lib.load().pixo_b200_png_filter(ctx.handle, d.ctypes.data, int(width), int(height),
                                          int(row_bytes), int(bytes_per_pixel),
                                          int(options.filter_strategy) | (OPTIMIZE_ALPHA if options.optimize_alpha else 0),
                                          out.ctypes.data,
                                          C.byref(ad) if with_adler else None)
    _lib.check(ctx.handle, rc)
    return (out, ad.value) if with_adler else out


def apply_filters(data, width, height, bytes_per_pixel, options: PngOptions, **kw):
    """filter::apply_filters (src/png/filter.rs:52)."""
    return apply_filters_with_row_bytes(data, width, height, int(width) * int(bytes_per_pixel),
                                        bytes_per_pixel, options, **kw)


def adler32(data, ctx: Context | None = None) -> int:
    """compress::adler32::adler32 (src/compress/adler32.rs:11)."""
    ctx = ctx or default_context()
    d = _as_u8(data)
    out = C.c_uint32()
    _lib.check(ctx.handle, _lib.load().pixo_b200_adler32(ctx.handle, d.ctypes.data if d.size else None,
                                                         d.size, C.byref(out)))
    return out.value


def apply_filters_rows_dev(d_rows, d_row_above, width, image_height, band_rows, row_bytes, bytes_per_pixel,
                           strategy: FilterStrategy, d_out, d_adler=None, optimize_alpha=False,
                           ctx: Context | None = None):
    """A band of rows of one image on the device (anything with .data_ptr()): see
    pixo_b200_png_filter_rows_dev.  Asynchronous on the context's stream."""
    ctx = ctx or default_context()
    p = lambda t: None if t is None else int(t.data_ptr())
    _lib.check(ctx.handle, _lib.load().pixo_b200_png_filter_rows_dev(
        ctx.handle, p(d_rows), p(d_row_above), int(width), int(image_height), int(band_rows), int(row_bytes),
        int(bytes_per_pixel), int(strategy) | (OPTIMIZE_ALPHA if optimize_alpha else 0), p(d_out), p(d_adler)))


def adler32_combine(adler_a: int, adler_b: int, len_b: int) -> int:
    """Adler-32 of A ++ B from the two checksums and len(B)."""
    return int(_lib.load().pixo_b200_adler32_combine(int(adler_a), int(adler_b), int(len_b)))
