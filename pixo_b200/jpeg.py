"""Mirror of pixo::jpeg for the accelerated path (src/jpeg/mod.rs:88-447).

`encode` / `encode_into` keep the reference's names, argument meaning and error behaviour;
the transform stages run on the GPU through libpixo_b200.so, the entropy stage on the host
inside the same library.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import enum

import numpy as np

from . import _lib
from .color import ColorType
from .context import Context, default_context


class Subsampling(enum.IntEnum):
    """pixo::jpeg::Subsampling (src/jpeg/mod.rs:96-102)."""
    S444 = 0
    S420 = 1


@dataclasses.dataclass
class JpegOptions:
    """pixo::jpeg::JpegOptions (src/jpeg/mod.rs:121-174): same fields, same defaults."""
    width: int = 0
    height: int = 0
    color_type: ColorType = ColorType.Rgb
    quality: int = 75
    subsampling: Subsampling = Subsampling.S444
    restart_interval: int | None = None
    optimize_huffman: bool = False
    progressive: bool = False
    trellis_quant: bool = False

    @classmethod
    def fast(cls, width, height, quality):
        return cls(width, height, ColorType.Rgb, quality, Subsampling.S444, None, False, False, False)

    @classmethod
    def balanced(cls, width, height, quality):
        return cls(width, height, ColorType.Rgb, quality, Subsampling.S444, None, True, False, False)

    @classmethod
    def max(cls, width, height, quality):
        return cls(width, height, ColorType.Rgb, quality, Subsampling.S420, None, True, True, True)

    @classmethod
    def from_preset(cls, width, height, quality, preset):
        return {0: cls.fast, 2: cls.max}.get(preset, cls.balanced)(width, height, quality)


def output_capacity(width: int, height: int) -> int:
    """Upper bound on a baseline JPEG's size: ~260 B per block worst case (63 sixteen-bit codes
    + ten-bit amplitudes), x2 for 0xFF stuffing, three full-resolution components, RSTn
    markers, headers."""
    nb = ((int(width) + 7) // 8) * ((int(height) + 7) // 8) * 3
    return nb * 600 + 4096


def _restart(options) -> int:
    """Some(0) is InvalidRestartInterval in the reference (src/jpeg/mod.rs:339-345); None -> 0."""
    if options.restart_interval is not None and int(options.restart_interval) == 0:
        raise _lib.PixoError(_lib.ERR_INVALID_RESTART, "Invalid restart interval 0")
    return int(options.restart_interval or 0)


def _as_u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    return np.frombuffer(data, dtype=np.uint8)


def quant_tables(quality: int):
    """QuantizationTables::with_quality (src/jpeg/quantize.rs:42-89): (lum_zz, chr_zz, lum, chr)."""
    lz = np.zeros(64, np.uint8); cz = np.zeros(64, np.uint8)
    ln = np.zeros(64, np.float32); cn = np.zeros(64, np.float32)
    _lib.load().pixo_b200_quant_tables(int(quality), lz.ctypes.data_as(_lib.u8p), cz.ctypes.data_as(_lib.u8p),
                                       ln.ctypes.data_as(_lib.f32p), cn.ctypes.data_as(_lib.f32p))
    return lz, cz, ln, cn


def block_counts(width, height, color_type=ColorType.Rgb, subsampling=Subsampling.S420):
    ny = C.c_size_t(); nc = C.c_size_t()
    _lib.check(None, _lib.load().pixo_b200_jpeg_block_counts(width, height, int(color_type), int(subsampling),
                                                             C.byref(ny), C.byref(nc)))
    return ny.value, nc.value


def compute_all_coefficients(data, width, height, color_type=ColorType.Rgb,
                             subsampling=Subsampling.S420, quality=None, lum_q=None, chr_q=None,
                             zigzag=False, histograms=False, ctx: Context | None = None):
    """compute_all_coefficients (src/jpeg/mod.rs:932-966) -> (y, cb, cr[, hist]) int16 [n,64]."""
    ctx = ctx or default_context()
    d = _as_u8(data)
    bpp = 1 if int(color_type) == ColorType.Gray else 3
    if d.size != int(width) * int(height) * bpp and width and height and int(color_type) in (0, 2):
        raise _lib.PixoError(_lib.ERR_INVALID_DATA_LENGTH,
                             f"Invalid data length: expected {int(width) * int(height) * bpp} bytes, got {d.size}")
    if lum_q is None:
        _, _, lum_q, chr_q = quant_tables(75 if quality is None else quality)
    lum_q = np.ascontiguousarray(lum_q, np.float32); chr_q = np.ascontiguousarray(chr_q, np.float32)
    ny, nc = block_counts(width, height, color_type, subsampling)
    y = np.empty((ny, 64), np.int16)
    cb = np.empty((max(nc, 1), 64), np.int16)
    cr = np.empty((max(nc, 1), 64), np.int16)
    hist = np.zeros(536, np.uint64) if histograms else None
    rc = _lib.load().pixo_b200_jpeg_coefficients(
        ctx.handle, d.ctypes.data, width, height, int(color_type), int(subsampling),
        lum_q.ctypes.data_as(_lib.f32p), chr_q.ctypes.data_as(_lib.f32p), y.ctypes.data,
        cb.ctypes.data, cr.ctypes.data, 1 if zigzag else 0, hist.ctypes.data if histograms else None)
    _lib.check(ctx.handle, rc)
    out = (y, cb[:nc], cr[:nc])
    return out + (hist,) if histograms else out


def encode_into(output: bytearray, data, options: JpegOptions, ctx: Context | None = None) -> None:
    """pixo::jpeg::encode_into (src/jpeg/mod.rs:328): clears and refills `output`."""
    ctx = ctx or default_context()
    d = _as_u8(data)
    cap = output_capacity(options.width, options.height)
    buf = np.empty(cap, np.uint8)
    n = C.c_size_t()
    restart = _restart(options)
    rc = _lib.load().pixo_b200_jpeg_encode(
        ctx.handle, d.ctypes.data, d.size, int(options.width), int(options.height),
        int(options.color_type), int(options.quality), int(options.subsampling),
        restart, int(bool(options.optimize_huffman)),
        int(bool(options.progressive)), int(bool(options.trellis_quant)), buf.ctypes.data, cap,
        C.byref(n))
    _lib.check(ctx.handle, rc)
    del output[:]
    output.extend(buf[: n.value].tobytes())


def encode(data, options: JpegOptions, ctx: Context | None = None) -> bytes:
    """pixo::jpeg::encode (src/jpeg/mod.rs:88)."""
    out = bytearray()
    encode_into(out, data, options, ctx)
    return bytes(out)


def encode_batch(frames: np.ndarray, options: JpegOptions, ctx: Context | None = None,
                 capacity_each: int | None = None) -> list[bytes]:
    """n frames of identical geometry ([n, h*w*bpp] uint8) -> n JPEG byte strings.
    capacity_each defaults to min(worst case, 2x the raw frame size)."""
    restart = _restart(options)
    ctx = ctx or default_context()
    f = np.ascontiguousarray(frames, np.uint8)
    n = f.shape[0]
    each = f.size // max(n, 1)
    cap = capacity_each or min(output_capacity(options.width, options.height), 2 * each + 4096)
    out = np.empty((n, cap), np.uint8)
    lens = (C.c_size_t * n)()
    rc = _lib.load().pixo_b200_jpeg_encode_batch(
        ctx.handle, f.ctypes.data, each, n, int(options.width), int(options.height),
        int(options.color_type), int(options.quality), int(options.subsampling),
        restart, int(bool(options.optimize_huffman)), out.ctypes.data,
        cap, lens)
    _lib.check(ctx.handle, rc)
    return [out[i, : lens[i]].tobytes() for i in range(n)]


def entropy_encode(y, cb, cr, options: JpegOptions, ctx: Context | None = None) -> bytes:
    """Host entropy stage on its own (no device needed)."""
    y = np.ascontiguousarray(y, np.int16)
    cb = np.ascontiguousarray(cb if len(cb) else np.zeros((1, 64)), np.int16)
    cr = np.ascontiguousarray(cr if len(cr) else np.zeros((1, 64)), np.int16)
    cap = output_capacity(options.width, options.height)
    buf = np.empty(cap, np.uint8)
    n = C.c_size_t()
    rc = _lib.load().pixo_b200_jpeg_entropy_encode(
        ctx.handle if ctx else None, y.ctypes.data, cb.ctypes.data, cr.ctypes.data,
        int(options.width), int(options.height), int(options.color_type), int(options.quality),
        int(options.subsampling), _restart(options),
        int(bool(options.optimize_huffman)), buf.ctypes.data, cap, C.byref(n))
    _lib.check(ctx.handle if ctx else None, rc)
    return buf[: n.value].tobytes()


def entropy_encode_dev(d_y, d_cb, d_cr, options: JpegOptions, ctx: Context | None = None) -> bytes:
    """Entropy-code coefficient arrays that live on the device (anything with .data_ptr(), e.g.
    int16 torch tensors in compute_all_coefficients' layout) into a complete JPEG: Huffman
    statistics, bit packing, stuffing and restart markers run on the GPU."""
    ctx = ctx or default_context()
    cap = output_capacity(options.width, options.height)
    buf = np.empty(cap, np.uint8)
    n = C.c_size_t()
    ptr = lambda t: None if t is None else int(t.data_ptr())
    rc = _lib.load().pixo_b200_jpeg_entropy_encode_dev(
        ctx.handle, ptr(d_y), ptr(d_cb), ptr(d_cr), int(options.width), int(options.height),
        int(options.color_type), int(options.quality), int(options.subsampling),
        _restart(options), int(bool(options.optimize_huffman)), buf.ctypes.data, cap,
        C.byref(n))
    _lib.check(ctx.handle, rc)
    return buf[: n.value].tobytes()
