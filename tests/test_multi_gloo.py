"""world_size-2 (and 3) gloo tests of the N>1 host path on CPU: frame sharding, MCU-row band planning,
and the DISTRIBUTED entropy stage of a tiled frame (parallel.encode_tiled): DC-predictor exchange,
optional histogram all-reduce, per-band raw coding, bit-offset exchange, splice, gather of the
finished scan bytes.  Band coefficients come from the oracle and the band stages run through the
library's host twins here (no GPU on the CPU box); on the GPU box the same collective code runs
with the CUDA band stages (tests/test_multi_gpu.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, w, h, q, ct, ss, opt, out_path):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    from pixo_b200 import parallel, synthetic
    ch = 1 if ct == 0 else 3
    frame = synthetic.noise(w, h, ch, 42)
    bands = parallel.plan_bands(w, h, world, gray=ct == 0, s420=ss == 1)
    b = bands[rank]
    px = np.ascontiguousarray(parallel.band_pixels(frame, w, h, ch, b)).reshape(-1)
    bh = b.px_row1 - b.px_row0
    z = np.zeros((0, 64), np.int16)
    y, cb, cr = po.jpeg_coefficients(px, w, bh, ct, ss, q) if bh else (z, z, z)
    assert y.shape[0] == b.y_blocks and (ct == 0 or cb.shape[0] == b.c_blocks)
    coder = parallel.HostBandCoder(y, cb if ct else z, cr if ct else z, w, max(bh, 1), ct, ss)
    jpg = parallel.encode_tiled(coder, w, h, ct, q, ss, opt, rank, world)
    # timing reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)]); dist.all_reduce(t, op=dist.ReduceOp.MAX); assert t.item() == world
    assert (jpg is not None) == (rank == 0)
    if rank == 0:
        open(out_path, "wb").write(jpg)
    dist.destroy_process_group()


@pytest.mark.parametrize("w,h,world,ct,ss,opt", [(200, 150, 2, 2, 1, False), (64, 16, 2, 2, 1, False),
                                                 (333, 517, 2, 2, 1, True), (100, 40, 3, 2, 0, True),
                                                 (77, 130, 2, 0, 0, False)])
def test_tiled_frame_distributed_entropy_is_byte_identical(po, tmp_path, w, h, world, ct, ss, opt):
    """(64,16): one MCU row over two ranks, (100,40,3): more ranks than..., i.e. EMPTY bands."""
    out = str(tmp_path / "tiled.jpg")
    mp.spawn(_worker, args=(world, _free_port(), w, h, 80, ct, ss, opt, out), nprocs=world, join=True)
    from pixo_b200 import synthetic
    ref = po.jpeg_encode(synthetic.noise(w, h, 1 if ct == 0 else 3, 42), w, h, ct, 80, ss, 0, opt)
    assert open(out, "rb").read() == ref


@pytest.mark.parametrize("world", [1, 3, 8, 13])
def test_band_stages_in_one_process(po, world):
    """encode_tiled_local (what bench.py's C4 line runs on one GPU) with the host twins: every
    bit phase and tail length occurs across these band counts; q=100 noise makes 0xFF bytes frequent,
    so bytes straddling bands get stuffed on either side."""
    from pixo_b200 import parallel, synthetic
    w, h = 144, 208
    for q, opt in ((100, False), (35, True)):
        frame = synthetic.noise(w, h, 3, 7)
        coders = []
        for b in parallel.plan_bands(w, h, world):
            bh = b.px_row1 - b.px_row0
            px = np.ascontiguousarray(parallel.band_pixels(frame, w, h, 3, b)).reshape(-1)
            z = np.zeros((0, 64), np.int16)
            y, cb, cr = po.jpeg_coefficients(px, w, bh, 2, 1, q) if bh else (z, z, z)
            coders.append(parallel.HostBandCoder(y, cb, cr, w, max(bh, 1), 2, 1))
        assert parallel.encode_tiled_local(coders, w, h, 2, q, 1, opt) == po.jpeg_encode(frame, w, h, 2, q, 1, 0, opt)


def test_splice_against_a_bitwise_model():
    """pixo_b200_jpeg_band_splice for every phase, with and without the final padding, against a
    bit-by-bit restatement of BitWriterMsb (src/bits.rs:216-272)."""
    import ctypes as C
    from pixo_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(5)
    for nbits in (0, 1, 7, 8, 9, 63, 64, 1000, 4099):
        raw = rng.integers(0, 256, (nbits + 7) // 8 + 1).astype(np.uint8)
        raw[rng.integers(0, raw.size, raw.size // 3)] = 0xFF
        bits = np.unpackbits(raw)[:nbits]
        for s in range(8):
            tail_in = int(rng.integers(0, 256)) & ((1 << s) - 1)
            tbits = np.concatenate([np.unpackbits(np.array([tail_in], np.uint8))[8 - s:] if s else np.zeros(0, np.uint8), bits])
            for last in (0, 1):
                t = tbits
                if last and t.size % 8:
                    t = np.concatenate([t, np.ones(8 - t.size % 8, np.uint8)])
                whole = np.packbits(t[: t.size // 8 * 8])
                want = bytearray()
                for v in whole:
                    want.append(int(v))
                    if v == 0xFF:
                        want.append(0)
                out = np.zeros(2 * raw.size + 16, np.uint8)
                n = C.c_size_t()
                rb = raw.copy(); rb[nbits // 8 + (1 if nbits % 8 else 0):] = 0
                if nbits % 8:
                    rb[nbits // 8] &= (0xFF << (8 - nbits % 8)) & 0xFF       # the coder zero-fills the last partial byte
                rc = lib.pixo_b200_jpeg_band_splice(rb.ctypes.data, nbits, 8 * 5 + s, tail_in, last, out.ctypes.data, out.size, C.byref(n))
                assert rc == 0 and bytes(out[: n.value]) == bytes(want), (nbits, s, last)


def test_sharding_plans_and_boundary_rules():
    from pixo_b200 import parallel
    for n, world in ((256, 8), (7, 4), (1, 8)):
        owned = [parallel.shard_frames(n, world, r) for r in range(world)]
        assert sorted(sum(owned, [])) == list(range(n))
        assert max(map(len, owned)) - min(map(len, owned)) <= 1
    bands = parallel.plan_bands(16384, 16384, 8)
    assert [b.mcu_row1 - b.mcu_row0 for b in bands] == [128] * 8           # 128 MCU rows / GPU
    assert sum(b.y_blocks for b in bands) == 4 * 1024 * 1024
    assert bands[-1].px_row1 == 16384 and all(b.px_row0 % 16 == 0 for b in bands)
    bands = parallel.plan_bands(100, 40, 8)    # 3 MCU rows over 8 ranks: empty bands allowed
    assert sum(b.mcu_row1 - b.mcu_row0 for b in bands) == 3
    assert parallel.plan_bands(100, 100, 2, gray=True)[0].c_blocks == 0
    # predictors and tails skip empty bands
    last = np.array([[5, 6, 7], [0, 0, 0], [9, 8, 7]])
    assert list(parallel.dc_seeds(last, [True, False, True], 2)) == [5, 6, 7]
    assert list(parallel.dc_seeds(last, [True, False, True], 0)) == [0, 0, 0]
    assert parallel.bit_offsets([13, 0, 20], [0b1010101, 0, 0b0000001], 2) == (13, 0b10101, True)
    assert parallel.bit_offsets([13, 0, 20], [0b1010101, 0, 1], 1) == (13, 0b10101, False)
    assert parallel.bit_offsets([16, 9], [3, 1], 1) == (16, 0, True)


def test_adler32_combine_matches_zlib():
    import zlib
    from pixo_b200 import parallel, png
    rng = np.random.default_rng(1)
    data = rng.integers(0, 256, 300001).astype(np.uint8).tobytes()
    cuts = [0, 1, 5553, 70000, 70000, 299999, len(data)]
    parts = [(zlib.adler32(data[a:b]), b - a) for a, b in zip(cuts, cuts[1:])]
    assert parallel.adler32_combine(parts) == zlib.adler32(data)
    acc = parts[0][0]
    for ad, ln in parts[1:]:
        acc = png.adler32_combine(acc, ad, ln)
    assert acc == zlib.adler32(data)
    big = (zlib.adler32(b"\xff" * 65521), 65521 * 70000)     # length far beyond the modulus
    assert parallel.adler32_combine([(1, 0), big]) == big[0]
