"""Multi-GPU host logic: one process per GPU (torch.distributed over NCCL/NVLink; gloo in the CPU
tests).  Nothing here moves pixels or coefficients between GPUs.

  * batches (BASELINE configs C2/C3/C5): frame i -> rank i % world (`shard_frames`).  Ranks never
    exchange data; only the timing reduction in bench.py is collective.
  * one gigapixel frame (config C4, `encode_tiled`): contiguous bands of MCU rows per rank.  Every
    MCU depends only on its own (clamped) pixels, so a band is transformed as an image of its own
    (bands start on MCU rows: no halo) - the analogue of the reference's
    compute_all_coefficients_parallel (src/jpeg/mod.rs:1137-1215).  The ENTROPY stage is distributed
    too: each rank Huffman-codes its own band into a raw bit string, two tiny all-gathers carry what
    crosses a band boundary (the DC predictors: 3 x i16; the band's bit count and last 7 bits), every
    rank then shifts its string to its bit offset in the frame's stream, completes the byte it
    shares with its predecessor, stuffs 0xFF and (last band) 1-pads, and only the finished scan
    bytes (~1/10 of the raw frame) are gathered to rank 0, which adds headers and EOI.  The file is
    byte-identical to a single-GPU (and to the reference's) encode.
  * PNG filter stream + Adler-32 of one image in row bands (`adler32_combine`): each band's
    checksum over its slice of the filtered stream, combined in band order.
"""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

from . import _lib


def shard_frames(n_frames: int, world: int, rank: int) -> list[int]:
    """Frame indices owned by `rank` (round robin keeps ragged batches balanced)."""
    return list(range(rank, n_frames, world))


@dataclasses.dataclass
class Band:
    rank: int
    mcu_row0: int
    mcu_row1: int      # exclusive
    px_row0: int
    px_row1: int       # exclusive, clipped to the image height
    y_blocks: int
    c_blocks: int


def plan_bands(width: int, height: int, world: int, gray: bool = False, s420: bool = True) -> list[Band]:
    """Contiguous MCU-row bands, as equal as possible; ranks beyond the number of MCU rows get
    empty bands."""
    mcu = 16 if (s420 and not gray) else 8
    mcus_x = (width + mcu - 1) // mcu
    mcus_y = (height + mcu - 1) // mcu
    ypm = 4 if (s420 and not gray) else 1
    bands = []
    for r in range(world):
        r0 = mcus_y * r // world
        r1 = mcus_y * (r + 1) // world
        n = (r1 - r0) * mcus_x
        bands.append(Band(r, r0, r1, r0 * mcu, min(r1 * mcu, height), n * ypm, 0 if gray else n))
    return bands


def band_pixels(frame: np.ndarray, width: int, height: int, bpp: int, band: Band) -> np.ndarray:
    """The rows of `frame` a band needs (a view): bands are MCU aligned, so no halo rows."""
    rows = np.asarray(frame, np.uint8).reshape(height, width * bpp)
    return rows[band.px_row0:band.px_row1]


# ---- what crosses a band boundary ----------------------------------------------------------------

def dc_seeds(last_dcs: np.ndarray, nonempty: list[bool], rank: int) -> np.ndarray:
    """DC predictors band `rank` starts from: the last DCs of the nearest non-empty band before it
    (0 for the first).  last_dcs: [world, 3]."""
    for r in range(rank - 1, -1, -1):
        if nonempty[r]:
            return np.asarray(last_dcs[r], np.int32).copy()
    return np.zeros(3, np.int32)


def bit_offsets(nbits: list[int], tails: list[int], rank: int) -> tuple[int, int, bool]:
    """(start_bit, tail_in, is_last) of band `rank` in the frame's stream.  start_bit = bits of all
    bands before it; tail_in = the last start_bit % 8 bits of the stream so far, taken from the
    nearest non-empty band (a non-empty band holds at least one MCU, i.e. more than 7 bits)."""
    start = int(sum(nbits[:rank]))
    s = start & 7
    tail_in = 0
    if s:
        for r in range(rank - 1, -1, -1):
            if nbits[r]:
                assert nbits[r] >= 7
                tail_in = int(tails[r]) & ((1 << s) - 1)
                break
    is_last = not any(nbits[r] for r in range(rank + 1, len(nbits)))
    return start, tail_in, is_last


class DeviceBandCoder:
    """The band stages on this rank's GPU (libpixo_b200: K3, k_huff<RAW>, k_splice_*)."""

    def __init__(self, ctx, d_y, d_cb, d_cr, width, band_height, color_type, subsampling, ny, nc):
        import torch
        self.ctx, self.lib, self.torch = ctx, _lib.load(), torch
        self.d_y, self.d_cb, self.d_cr = d_y, d_cb, d_cr
        self.geo = (int(width), int(band_height), int(color_type), int(subsampling))
        self.ny, self.nc = int(ny), int(nc)
        self.dev = d_y.device
        self.raw = None
        self.out = None

    def _p(self, t):
        return None if t is None else int(t.data_ptr())

    def last_dc(self) -> np.ndarray:
        out = (C.c_int32 * 3)()
        _lib.check(self.ctx.handle, self.lib.pixo_b200_jpeg_band_last_dc(
            self.ctx.handle, self._p(self.d_y), self._p(self.d_cb), self._p(self.d_cr), self.ny, self.nc, out))
        return np.array(out[:], np.int32)

    def histogram(self, seed: np.ndarray):
        """536 counters of this band (torch int64 on the device, ready for all_reduce)."""
        # (the library clears the counters itself, on ITS stream: a torch-side fill could land after it)
        hist = self.torch.empty(536, dtype=self.torch.int64, device=self.dev)
        if not self.ny:
            hist.zero_()
        if self.ny:
            s = (C.c_int32 * 3)(*[int(v) for v in seed])
            _lib.check(self.ctx.handle, self.lib.pixo_b200_jpeg_band_histogram_dev(
                self.ctx.handle, self._p(self.d_y), self._p(self.d_cb), self._p(self.d_cr), *self.geo, s,
                int(hist.data_ptr())))
            self.ctx.sync()
        return hist

    def entropy(self, seed: np.ndarray, hist: np.ndarray | None) -> tuple[int, int]:
        if not self.ny:
            return 0, 0
        w, bh = self.geo[0], self.geo[1]
        # as many bytes as the band's pixels + 1 MiB: lets a long band be coded in segments (short
        # look-back chains), whose raw strings need more room than the finished JPEG; grown on demand
        cap = (w * bh * 3 + (1 << 20)) // 16 * 16
        s = (C.c_int32 * 3)(*[int(v) for v in seed])
        hp = None if hist is None else np.ascontiguousarray(hist, np.uint64).ctypes.data_as(_lib.u64p)
        for _ in range(2):
            if self.raw is None or self.raw.numel() < cap:
                self.raw = self.torch.empty(cap, dtype=self.torch.uint8, device=self.dev)
            cap = self.raw.numel() // 16 * 16
            nbits, tail = C.c_uint64(), C.c_uint32()
            rc = self.lib.pixo_b200_jpeg_band_entropy_dev(
                self.ctx.handle, self._p(self.d_y), self._p(self.d_cb), self._p(self.d_cr), *self.geo, s, hp,
                int(self.raw.data_ptr()), cap, C.byref(nbits), C.byref(tail))
            if rc == _lib.ERR_OUTPUT_TOO_SMALL:
                cap = ((nbits.value + 7) // 8 + 4096) // 16 * 16
                continue
            _lib.check(self.ctx.handle, rc)
            return int(nbits.value), int(tail.value)
        _lib.check(self.ctx.handle, rc)

    def splice(self, nbits, start_bit, tail_in, is_last):
        """-> uint8 tensor (device) of this band's finished scan bytes."""
        cap = (nbits + 7) // 8 * 2 + 64     # worst case: every byte stuffed
        if self.out is None or self.out.numel() < cap:
            self.out = self.torch.empty(cap, dtype=self.torch.uint8, device=self.dev)
        n = C.c_uint64()
        _lib.check(self.ctx.handle, self.lib.pixo_b200_jpeg_band_splice_dev(
            self.ctx.handle, self._p(self.raw) if self.raw is not None else None, nbits, start_bit, tail_in,
            int(is_last), int(self.out.data_ptr()), self.out.numel(), C.byref(n)))
        return self.out[: n.value]


class HostBandCoder:
    """The same stages from host arrays through the library's host twins (no device): the CPU-only
    world_size-2 tests run the collective logic with it."""

    def __init__(self, y, cb, cr, width, band_height, color_type, subsampling):
        import torch
        self.lib, self.torch = _lib.load(), torch
        self.y = np.ascontiguousarray(y, np.int16).reshape(-1, 64)
        self.cb = np.ascontiguousarray(cb, np.int16).reshape(-1, 64)
        self.cr = np.ascontiguousarray(cr, np.int16).reshape(-1, 64)
        self.geo = (int(width), int(band_height), int(color_type), int(subsampling))
        self.ny = self.y.shape[0]
        self.dev = torch.device("cpu")
        self.raw = None

    def _ptrs(self):
        z = np.zeros((1, 64), np.int16)
        cb = self.cb if len(self.cb) else z
        cr = self.cr if len(self.cr) else z
        self._keep = (cb, cr)
        return self.y.ctypes.data, cb.ctypes.data, cr.ctypes.data

    def last_dc(self) -> np.ndarray:
        v = [int(a[-1, 0]) if len(a) else 0 for a in (self.y, self.cb, self.cr)]
        return np.array(v, np.int32)

    def histogram(self, seed):
        hist = np.zeros(536, np.uint64)
        if self.ny:
            s = (C.c_int32 * 3)(*[int(v) for v in seed])
            _lib.check(None, self.lib.pixo_b200_jpeg_band_histogram(*self._ptrs(), *self.geo, s,
                                                                    hist.ctypes.data_as(_lib.u64p)))
        return self.torch.from_numpy(hist.astype(np.int64))

    def entropy(self, seed, hist):
        if not self.ny:
            return 0, 0
        cap = self.y.size * 4 + self.cb.size * 8 + 4096
        self.raw = np.zeros(cap, np.uint8)
        s = (C.c_int32 * 3)(*[int(v) for v in seed])
        hp = None if hist is None else np.ascontiguousarray(hist, np.uint64).ctypes.data_as(_lib.u64p)
        nbits, tail = C.c_uint64(), C.c_uint32()
        _lib.check(None, self.lib.pixo_b200_jpeg_band_entropy(*self._ptrs(), *self.geo, s, hp, self.raw.ctypes.data,
                                                              cap, C.byref(nbits), C.byref(tail)))
        return int(nbits.value), int(tail.value)

    def splice(self, nbits, start_bit, tail_in, is_last):
        cap = (nbits + 7) // 8 * 2 + 64
        out = np.zeros(cap, np.uint8)
        n = C.c_size_t()
        _lib.check(None, self.lib.pixo_b200_jpeg_band_splice(self.raw.ctypes.data if self.raw is not None else None,
                                                             nbits, start_bit, tail_in, int(is_last), out.ctypes.data,
                                                             cap, C.byref(n)))
        return self.torch.from_numpy(out[: n.value])


def write_headers(width, height, color_type, quality, subsampling, hist=None) -> bytes:
    buf = np.zeros(2048, np.uint8)
    n = C.c_size_t()
    hp = None if hist is None else np.ascontiguousarray(hist, np.uint64).ctypes.data_as(_lib.u64p)
    _lib.check(None, _lib.load().pixo_b200_jpeg_write_headers(int(width), int(height), int(color_type), int(quality),
                                                              int(subsampling), 0, hp, buf.ctypes.data, buf.size,
                                                              C.byref(n)))
    return buf[: n.value].tobytes()


def encode_tiled(coder, width: int, height: int, color_type: int, quality: int, subsampling: int,
                 optimize_huffman: bool, rank: int, world: int, dst: int = 0):
    """Distributed entropy stage of one tiled frame: the complete JPEG on `dst`, None elsewhere."""
    parts, hist = tiled_scan_parts(coder, optimize_huffman, rank, world, dst)
    if rank != dst:
        return None
    return assemble_tiled(parts, hist, width, height, color_type, quality, subsampling)


def assemble_tiled(parts, hist, width, height, color_type, quality, subsampling) -> bytes:
    """Host side of the last step: headers + the bands' scan bytes (device or host tensors) + EOI."""
    import torch
    scan = torch.cat([p.cpu() for p in parts]).numpy().tobytes() if parts else b""
    return write_headers(width, height, color_type, quality, subsampling, hist) + scan + b"\xff\xd9"


def tiled_scan_parts(coder, optimize_huffman: bool, rank: int, world: int, dst: int = 0):
    """The device part of `encode_tiled`.  `coder` holds this rank's band (coefficients already
    computed).  Returns (the bands' finished scan bytes in band order as tensors on dst's device -
    None on the other ranks -, the summed histogram or None).  Collectives: all_gather of 4 ints (DC
    predictors), [all_reduce of 536 counters], all_gather of 2 ints (bits, tail), all_gather of 1 int
    (byte counts), gather of the scan bytes (only dst receives)."""
    import torch
    import torch.distributed as dist
    dev = coder.dev

    def all_gather_ints(vals):
        t = torch.tensor(vals, dtype=torch.int64, device=dev)
        if world == 1:
            return t.reshape(1, -1).cpu().numpy()
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return torch.stack(out).cpu().numpy()

    ld = coder.last_dc()
    g = all_gather_ints([int(ld[0]), int(ld[1]), int(ld[2]), int(coder.ny > 0)])
    seed = dc_seeds(g[:, :3], [bool(v) for v in g[:, 3]], rank)
    hist = None
    if optimize_huffman:
        h = coder.histogram(seed)
        if world > 1:
            dist.all_reduce(h, op=dist.ReduceOp.SUM)
        hist = h.cpu().numpy().astype(np.uint64)
    nbits, tail = coder.entropy(seed, hist)
    g = all_gather_ints([nbits, tail])
    start, tail_in, is_last = bit_offsets([int(v) for v in g[:, 0]], [int(v) for v in g[:, 1]], rank)
    if nbits == 0:
        is_last = False          # an empty band owns nothing; the last NON-empty band pads
    body = coder.splice(nbits, start, tail_in, is_last) if nbits else torch.empty(0, dtype=torch.uint8, device=dev)
    sizes = all_gather_ints([int(body.numel())])[:, 0]
    if world == 1:
        parts = [body]
    else:
        mx = int(max(sizes.max(), 1))
        pad = torch.zeros(mx, dtype=torch.uint8, device=dev)
        pad[: body.numel()] = body
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, bufs, dst=dst)      # every rank sends its (padded) bytes, only dst receives
        parts = [bufs[r][: int(sizes[r])] for r in range(world)] if rank == dst else None
    return (parts if rank == dst else None), hist


class ThreadComm:
    """all_gather / gather among `world` THREADS of one process (one band per thread, every band with
    a context of its own on the same GPU): lets the multi-rank flow run - collectives included - where
    there is only one device.  Device work is synchronised before tensors change hands."""

    def __init__(self, world: int):
        import threading
        self.world = world
        self.slots = [None] * world
        self.barrier = threading.Barrier(world)

    def all_gather(self, rank, t):
        import torch
        torch.cuda.current_stream().synchronize()
        self.slots[rank] = t
        self.barrier.wait()
        out = torch.stack([x.clone() for x in self.slots])
        torch.cuda.current_stream().synchronize()
        self.barrier.wait()
        return out

    def gather(self, rank, t, dst):
        allv = self.all_gather(rank, t)
        return [allv[r] for r in range(self.world)] if rank == dst else None


class _DistComm:
    def __init__(self, world):
        self.world = world

    def all_gather(self, rank, t):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return t.reshape(1, *t.shape)
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t)
        return torch.stack(out)

    def gather(self, rank, t, dst):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return [t]
        bufs = [torch.empty_like(t) for _ in range(self.world)] if rank == dst else None
        dist.gather(t, bufs, dst=dst)
        return bufs


def tiled_scan_parts_async(coder, nonempty: list[bool], rank: int, world: int, dst: int = 0, comm=None):
    """`tiled_scan_parts` without host round trips: every value that crosses a band boundary stays
    on the device (pixo_b200_jpeg_band_*_async), the collectives are queued on the same stream as the
    kernels, and the host waits once - for the byte counts it needs to size the final gather.
    Standard Huffman tables only (optimised tables need the all-reduced statistics on the host).
    `coder` is a DeviceBandCoder whose context runs on torch's current stream; nonempty[r] = band r
    holds at least one MCU (known from the band plan)."""
    import torch
    dev, lib, ctx = coder.dev, coder.lib, coder.ctx
    i64 = torch.int64
    comm = comm or _DistComm(world)
    all_gather = lambda t: comm.all_gather(rank, t)

    prev = next((r for r in range(rank - 1, -1, -1) if nonempty[r]), None)
    is_last = coder.ny > 0 and not any(nonempty[rank + 1:])
    # 1. DC predictors: the last DC of every band's three arrays
    if coder.ny:
        ld = torch.stack([coder.d_y[coder.ny - 1, 0], coder.d_cb[coder.nc - 1, 0] if coder.nc else coder.d_y[0, 0] * 0,
                          coder.d_cr[coder.nc - 1, 0] if coder.nc else coder.d_y[0, 0] * 0]).to(torch.int32)
    else:
        ld = torch.zeros(3, dtype=torch.int32, device=dev)
    g = all_gather(ld)
    seed = (g[prev] if prev is not None else torch.zeros(3, dtype=torch.int32, device=dev)).contiguous()
    # 2. this band's code as a raw bit string
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    bits_tail = torch.zeros(2, dtype=i64, device=dev)
    if coder.ny:
        w, bh = coder.geo[0], coder.geo[1]
        cap = (w * bh * 3 + (1 << 20)) // 16 * 16
        if coder.raw is None or coder.raw.numel() < cap:
            coder.raw = torch.empty(cap, dtype=torch.uint8, device=dev)
        _lib.check(ctx.handle, lib.pixo_b200_jpeg_band_entropy_dev_async(
            ctx.handle, coder._p(coder.d_y), coder._p(coder.d_cb), coder._p(coder.d_cr), *coder.geo, int(seed.data_ptr()),
            None, int(coder.raw.data_ptr()), coder.raw.numel() // 16 * 16, int(bits_tail.data_ptr()), int(flags.data_ptr())))
    # 3. bit offsets: bits of all bands before this one; the previous non-empty band's last 7 bits
    bt = all_gather(bits_tail)
    start = bt[:rank, 0].sum() if rank else torch.zeros((), dtype=i64, device=dev)
    tail_prev = bt[prev, 1] if prev is not None else torch.zeros((), dtype=i64, device=dev)
    offset = torch.stack([start, tail_prev, torch.full((), int(is_last), dtype=i64, device=dev)]).contiguous()
    # 4. splice
    out_len = torch.zeros(1, dtype=i64, device=dev)
    if coder.ny:
        cap = coder.raw.numel() * 2 + 64
        if coder.out is None or coder.out.numel() < cap:
            coder.out = torch.empty(cap, dtype=torch.uint8, device=dev)
        _lib.check(ctx.handle, lib.pixo_b200_jpeg_band_splice_dev_async(
            ctx.handle, int(coder.raw.data_ptr()), int(offset.data_ptr()), int(coder.out.data_ptr()), coder.out.numel(),
            int(out_len.data_ptr()), int(flags.data_ptr())))
    # 5. the one host wait: byte counts (and flags) of all bands
    info = all_gather(torch.cat([out_len, flags.to(i64)])).cpu().numpy()
    if int(info[:, 1].max()):
        raise _lib.PixoError(_lib.ERR_CUDA, f"band entropy stage reported flags {info[:, 1].tolist()}")
    sizes = info[:, 0]
    body = coder.out[: int(sizes[rank])] if coder.ny else torch.empty(0, dtype=torch.uint8, device=dev)
    if world == 1:
        return [body], None
    mx = int(max(sizes.max(), 1))
    pad = torch.zeros(mx, dtype=torch.uint8, device=dev)
    pad[: body.numel()] = body
    bufs = comm.gather(rank, pad, dst)
    return ([bufs[r][: int(sizes[r])] for r in range(world)] if rank == dst else None), None


def encode_tiled_local(coders: list, width: int, height: int, color_type: int, quality: int, subsampling: int,
                       optimize_huffman: bool = False) -> bytes:
    """The same stage sequence with every band in THIS process (one GPU context, or the host
    twins): what `encode_tiled` does across ranks, minus the collectives.  Used at world size 1
    (bench.py's C4 line on one GPU, the single-device tests)."""
    parts, hist = tiled_scan_parts_local(coders, optimize_huffman)
    return assemble_tiled(parts, hist, width, height, color_type, quality, subsampling)


def tiled_scan_parts_local(coders: list, optimize_huffman: bool = False):
    import torch
    world = len(coders)
    last = np.stack([c.last_dc() for c in coders])
    nonempty = [c.ny > 0 for c in coders]
    seeds = [dc_seeds(last, nonempty, r) for r in range(world)]
    hist = None
    if optimize_huffman:
        hist = sum(c.histogram(seeds[r]).cpu().numpy().astype(np.uint64) for r, c in enumerate(coders))
    coded = [c.entropy(seeds[r], hist) for r, c in enumerate(coders)]
    nbits = [n for n, _ in coded]
    tails = [t for _, t in coded]
    parts = []
    for r, c in enumerate(coders):
        if not nbits[r]:
            continue
        start, tail_in, is_last = bit_offsets(nbits, tails, r)
        parts.append(c.splice(nbits[r], start, tail_in, is_last))
    return parts, hist


# ---- Adler-32 of a stream held in pieces (PNG filter stage in row bands) ----------------------------

ADLER_MOD = 65521


def adler32_combine(parts: list[tuple[int, int]]) -> int:
    """Combine per-piece Adler-32 values, in stream order.  parts: (adler32 of the piece computed
    from the initial state s1=1,s2=0 as compress::adler32::adler32 does - src/compress/adler32.rs:26-47 -,
    piece length).  For pieces A then B:  s1 = s1A + s1B - 1,  s2 = s2A + s2B + lenB * (s1A - 1)  (mod 65521)."""
    s1, s2 = 1, 0
    for ad, ln in parts:
        b1, b2 = ad & 0xFFFF, (ad >> 16) & 0xFFFF
        s2 = (s2 + b2 + (ln % ADLER_MOD) * (s1 - 1)) % ADLER_MOD
        s1 = (s1 + b1 - 1) % ADLER_MOD
    return (s2 << 16) | s1
