"""ctypes binding of the CPU oracle (oracle/libpixo_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (pixo_b200) never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpixo_oracle.so")

GRAY, GRAY_ALPHA, RGB, RGBA = 0, 1, 2, 3
S444, S420 = 0, 1
F_NONE, F_SUB, F_UP, F_AVERAGE, F_PAETH, F_MINSUM, F_ADAPTIVE, F_ADAPTIVE_FAST, F_BIGRAMS = range(9)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pixo_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpixo_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None
u8p = C.POINTER(C.c_uint8)
i16p = C.POINTER(C.c_int16)
f32p = C.POINTER(C.c_float)
u64p = C.POINTER(C.c_uint64)


class HuffTables(C.Structure):
    _fields_ = [("bits", (C.c_uint8 * 16) * 4), ("vals", (C.c_uint8 * 256) * 4),
                ("nvals", C.c_int * 4), ("code", (C.c_uint16 * 256) * 4),
                ("len", (C.c_uint8 * 256) * 4)]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.po_jpeg_encode.restype = C.c_long
        L.po_jpeg_encode.argtypes = [u8p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                     C.c_int, C.c_uint32, C.c_int, u8p, C.c_size_t]
        L.po_jpeg_encode_from_coefficients.restype = C.c_long
        L.po_jpeg_encode_from_coefficients.argtypes = [
            i16p, i16p, i16p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_uint32,
            C.c_int, u8p, C.c_size_t]
        L.po_jpeg_coefficients.restype = None
        L.po_jpeg_coefficients.argtypes = [u8p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, f32p,
                                           f32p, i16p, i16p, i16p, C.c_uint32, C.c_uint32]
        L.po_jpeg_block_counts.restype = None
        L.po_jpeg_block_counts.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                           C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.po_jpeg_histograms.restype = None
        L.po_jpeg_histograms.argtypes = [i16p, i16p, i16p, C.c_uint32, C.c_uint32, C.c_int,
                                         C.c_int, C.c_uint32, u64p]
        L.po_quant_tables.restype = None
        L.po_quant_tables.argtypes = [C.c_int, u8p, u8p, f32p, f32p]
        L.po_dct_2d.restype = None
        L.po_dct_2d.argtypes = [f32p, f32p]
        L.po_quantize_block.restype = None
        L.po_quantize_block.argtypes = [f32p, f32p, i16p]
        L.po_zigzag_reorder.restype = None
        L.po_zigzag_reorder.argtypes = [i16p, i16p]
        L.po_rgb_to_ycbcr.restype = None
        L.po_rgb_to_ycbcr.argtypes = [C.c_uint8, C.c_uint8, C.c_uint8, u8p]
        L.po_block_symbols.restype = C.c_int
        L.po_block_symbols.argtypes = [i16p, C.c_int16, u8p, C.POINTER(C.c_uint16), u8p, i16p]
        L.po_huff_standard.restype = None
        L.po_huff_standard.argtypes = [C.POINTER(HuffTables)]
        L.po_huff_optimized.restype = C.c_int
        L.po_huff_optimized.argtypes = [u64p, C.c_int, C.POINTER(HuffTables)]
        for name in ("po_filter_sub",):
            getattr(L, name).restype = None
        L.po_filter_sub.argtypes = [u8p, C.c_size_t, C.c_size_t, u8p]
        L.po_filter_up.restype = None
        L.po_filter_up.argtypes = [u8p, u8p, C.c_size_t, u8p]
        L.po_filter_average.restype = None
        L.po_filter_average.argtypes = [u8p, u8p, C.c_size_t, C.c_size_t, u8p]
        L.po_filter_paeth.restype = None
        L.po_filter_paeth.argtypes = [u8p, u8p, C.c_size_t, C.c_size_t, u8p]
        L.po_paeth_predictor.restype = C.c_uint8
        L.po_paeth_predictor.argtypes = [C.c_uint8, C.c_uint8, C.c_uint8]
        L.po_score_filter.restype = C.c_uint64
        L.po_score_filter.argtypes = [u8p, C.c_size_t]
        L.po_score_bigrams.restype = C.c_size_t
        L.po_score_bigrams.argtypes = [u8p, C.c_size_t]
        L.po_apply_filters.restype = None
        L.po_apply_filters.argtypes = [u8p, C.c_uint32, C.c_uint32, C.c_size_t, C.c_size_t,
                                       C.c_int, C.c_int, u8p, C.c_uint32, C.c_uint32]
        L.po_optimize_alpha.restype = None
        L.po_optimize_alpha.argtypes = [u8p, C.c_size_t, C.c_int]
        L.po_adler32.restype = C.c_uint32
        L.po_adler32.argtypes = [u8p, C.c_size_t]
        L.po_crc32.restype = C.c_uint32
        L.po_crc32.argtypes = [u8p, C.c_size_t]
        L.po_gen_gradient_rgb.restype = None
        L.po_gen_gradient_rgb.argtypes = [C.c_uint32, C.c_uint32, u8p]
        L.po_gen_noise.restype = None
        L.po_gen_noise.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u8p]
        _lib = L
    return _lib


def _u8(a):
    return a.ctypes.data_as(u8p)


def _as_u8(data) -> np.ndarray:
    a = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray))
                             else np.asarray(data, dtype=np.uint8))
    return a.reshape(-1)


# ---- JPEG ------------------------------------------------------------------------------
def quant_tables(quality: int):
    lz = np.zeros(64, np.uint8); cz = np.zeros(64, np.uint8)
    ln = np.zeros(64, np.float32); cn = np.zeros(64, np.float32)
    lib().po_quant_tables(quality, _u8(lz), _u8(cz), ln.ctypes.data_as(f32p), cn.ctypes.data_as(f32p))
    return lz, cz, ln, cn


def block_counts(w, h, color_type=RGB, subsampling=S420):
    ny = C.c_size_t(); nc = C.c_size_t()
    lib().po_jpeg_block_counts(w, h, color_type, subsampling, C.byref(ny), C.byref(nc))
    return ny.value, nc.value


def jpeg_coefficients(data, w, h, color_type=RGB, subsampling=S420, quality=80,
                      lum_q=None, chr_q=None, mcu_rows=(0, 0)):
    d = _as_u8(data)
    if lum_q is None:
        _, _, lum_q, chr_q = quant_tables(quality)
    lum_q = np.ascontiguousarray(lum_q, np.float32); chr_q = np.ascontiguousarray(chr_q, np.float32)
    ny, nc = block_counts(w, h, color_type, subsampling)
    y = np.zeros((ny, 64), np.int16)
    cb = np.zeros((max(nc, 1), 64), np.int16)
    cr = np.zeros((max(nc, 1), 64), np.int16)
    lib().po_jpeg_coefficients(_u8(d), w, h, color_type, subsampling,
                               lum_q.ctypes.data_as(f32p), chr_q.ctypes.data_as(f32p),
                               y.ctypes.data_as(i16p), cb.ctypes.data_as(i16p),
                               cr.ctypes.data_as(i16p), mcu_rows[0], mcu_rows[1])
    return y, cb[:nc], cr[:nc]


def jpeg_histograms(y, cb, cr, w, h, color_type=RGB, subsampling=S420, restart_interval=0):
    hist = np.zeros(536, np.uint64)
    cb = cb if len(cb) else np.zeros((1, 64), np.int16)
    cr = cr if len(cr) else np.zeros((1, 64), np.int16)
    lib().po_jpeg_histograms(np.ascontiguousarray(y).ctypes.data_as(i16p),
                             np.ascontiguousarray(cb).ctypes.data_as(i16p),
                             np.ascontiguousarray(cr).ctypes.data_as(i16p),
                             w, h, color_type, subsampling, restart_interval,
                             hist.ctypes.data_as(u64p))
    return hist


def _jpeg_cap(w, h) -> int:
    # worst case: every block codes 63 sixteen-bit symbols with ten-bit amplitudes (~260 B),
    # three components at full resolution, stuffing can double it, plus headers and RSTn
    nb = ((int(w) + 7) // 8) * ((int(h) + 7) // 8) * 3
    return nb * 600 + 4096


def jpeg_encode(data, w, h, color_type=RGB, quality=80, subsampling=S420, restart_interval=0,
                optimize_huffman=False, out: np.ndarray | None = None) -> bytes:
    d = _as_u8(data)
    cap = _jpeg_cap(w, h)
    if out is None or out.size < cap:
        out = np.empty(cap, np.uint8)
    n = lib().po_jpeg_encode(_u8(d), d.size, w, h, color_type, quality, subsampling,
                             restart_interval or 0, int(optimize_huffman), _u8(out), out.size)
    if n < 0:
        raise ValueError(f"po_jpeg_encode error {n}")
    return out[:n].tobytes()


def jpeg_encode_from_coefficients(y, cb, cr, w, h, color_type=RGB, quality=80, subsampling=S420,
                                  restart_interval=0, optimize_huffman=False) -> bytes:
    cap = _jpeg_cap(w, h)
    out = np.empty(cap, np.uint8)
    cb = cb if len(cb) else np.zeros((1, 64), np.int16)
    cr = cr if len(cr) else np.zeros((1, 64), np.int16)
    n = lib().po_jpeg_encode_from_coefficients(
        np.ascontiguousarray(y).ctypes.data_as(i16p), np.ascontiguousarray(cb).ctypes.data_as(i16p),
        np.ascontiguousarray(cr).ctypes.data_as(i16p), w, h, color_type, quality, subsampling,
        restart_interval or 0, int(optimize_huffman), _u8(out), out.size)
    if n < 0:
        raise ValueError(f"po_jpeg_encode_from_coefficients error {n}")
    return out[:n].tobytes()


def dct_2d(block) -> np.ndarray:
    b = np.ascontiguousarray(block, np.float32).reshape(64)
    o = np.zeros(64, np.float32)
    lib().po_dct_2d(b.ctypes.data_as(f32p), o.ctypes.data_as(f32p))
    return o


def quantize_block(dct, q) -> np.ndarray:
    d = np.ascontiguousarray(dct, np.float32).reshape(64)
    qq = np.ascontiguousarray(q, np.float32).reshape(64)
    o = np.zeros(64, np.int16)
    lib().po_quantize_block(d.ctypes.data_as(f32p), qq.ctypes.data_as(f32p), o.ctypes.data_as(i16p))
    return o


def zigzag_reorder(block) -> np.ndarray:
    b = np.ascontiguousarray(block, np.int16).reshape(64)
    o = np.zeros(64, np.int16)
    lib().po_zigzag_reorder(b.ctypes.data_as(i16p), o.ctypes.data_as(i16p))
    return o


def rgb_to_ycbcr(r, g, b):
    o = np.zeros(3, np.uint8)
    lib().po_rgb_to_ycbcr(r, g, b, _u8(o))
    return int(o[0]), int(o[1]), int(o[2])


def block_symbols(nat, prev_dc=0):
    b = np.ascontiguousarray(nat, np.int16).reshape(64)
    rs = np.zeros(65, np.uint8); amp = np.zeros(65, np.uint16); nb = np.zeros(65, np.uint8)
    dc = C.c_int16()
    n = lib().po_block_symbols(b.ctypes.data_as(i16p), prev_dc, _u8(rs),
                               amp.ctypes.data_as(C.POINTER(C.c_uint16)), _u8(nb), C.byref(dc))
    return rs[:n].copy(), amp[:n].copy(), nb[:n].copy(), dc.value


# ---- PNG -------------------------------------------------------------------------------
def _filt(fn, row, prev, bpp):
    r = _as_u8(row); o = np.zeros(max(r.size, 1), np.uint8)
    if fn == "sub":
        lib().po_filter_sub(_u8(r), r.size, bpp, _u8(o))
    else:
        p = _as_u8(prev)
        if fn == "up":
            lib().po_filter_up(_u8(r), _u8(p), r.size, _u8(o))
        elif fn == "average":
            lib().po_filter_average(_u8(r), _u8(p), r.size, bpp, _u8(o))
        else:
            lib().po_filter_paeth(_u8(r), _u8(p), r.size, bpp, _u8(o))
    return o[:r.size]


def filter_sub(row, bpp): return _filt("sub", row, None, bpp)
def filter_up(row, prev): return _filt("up", row, prev, 1)
def filter_average(row, prev, bpp): return _filt("average", row, prev, bpp)
def filter_paeth(row, prev, bpp): return _filt("paeth", row, prev, bpp)
def paeth_predictor(a, b, c): return int(lib().po_paeth_predictor(a, b, c))


def score_filter(f) -> int:
    a = _as_u8(f)
    return int(lib().po_score_filter(_u8(a), a.size))


def score_bigrams(f) -> int:
    a = _as_u8(f)
    return int(lib().po_score_bigrams(_u8(a), a.size))


def apply_filters(data, width, height, bpp, strategy, row_bytes=None, parallel_feature=True,
                  rows=(0, 0)) -> np.ndarray:
    d = _as_u8(data)
    if row_bytes is None:
        row_bytes = width * bpp
    out = np.zeros(height * (row_bytes + 1), np.uint8)
    lib().po_apply_filters(_u8(d), width, height, row_bytes, bpp, strategy, int(parallel_feature),
                           _u8(out), rows[0], rows[1])
    return out


def optimize_alpha(data, color_type) -> np.ndarray:
    """maybe_optimize_alpha(data, color_type, true) (src/png/mod.rs:633-671); returns a copy."""
    a = _as_u8(data).copy()
    lib().po_optimize_alpha(_u8(a), a.size, int(color_type))
    return a


def adler32(data) -> int:
    a = _as_u8(data)
    return int(lib().po_adler32(_u8(a), a.size))


def crc32(data) -> int:
    a = _as_u8(data)
    return int(lib().po_crc32(_u8(a), a.size))


# ---- inputs ----------------------------------------------------------------------------
def gen_gradient_rgb(w, h) -> np.ndarray:
    o = np.zeros(w * h * 3, np.uint8)
    lib().po_gen_gradient_rgb(w, h, _u8(o))
    return o


def gen_noise(w, h, channels=3, seed=42) -> np.ndarray:
    o = np.zeros(w * h * channels, np.uint8)
    lib().po_gen_noise(w, h, channels, seed, _u8(o))
    return o
