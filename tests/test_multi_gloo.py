"""world_size-2 gloo test of the N>1 host path on CPU: frame sharding, MCU-row band planning,
coefficient gather in band order, and entropy coding of the gathered arrays by the product's
host coder.  Band coefficients come from the oracle here (no GPU on the CPU box); on the GPU box
the same flow runs with the CUDA transform (tests/test_multi_gpu.py)."""
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


def _worker(rank, world, port, w, h, q, out_path):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    from pixo_b200 import ColorType, parallel, synthetic
    from pixo_b200.jpeg import JpegOptions, Subsampling, entropy_encode
    frame = synthetic.noise(w, h, 3, 42)
    bands = parallel.plan_bands(w, h, world)
    b = bands[rank]
    px = np.ascontiguousarray(parallel.band_pixels(frame, w, h, 3, b)).reshape(-1)
    bh = b.px_row1 - b.px_row0
    y, cb, cr = po.jpeg_coefficients(px, w, bh, po.RGB, po.S420, q) if bh else (np.zeros((0, 64), np.int16),) * 3
    assert y.shape[0] == b.y_blocks and cb.shape[0] == b.c_blocks
    gy, gcb, gcr = parallel.gather_coefficients(torch.from_numpy(y), torch.from_numpy(cb), torch.from_numpy(cr),
                                                bands, rank, world)
    # timing reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)]); dist.all_reduce(t, op=dist.ReduceOp.MAX); assert t.item() == world
    if rank == 0:
        jpg = entropy_encode(gy.numpy(), gcb.numpy(), gcr.numpy(), JpegOptions(w, h, ColorType.Rgb, q, Subsampling.S420))
        open(out_path, "wb").write(jpg)
    dist.destroy_process_group()


@pytest.mark.parametrize("w,h", [(200, 150), (64, 16), (333, 517)])
def test_two_rank_tiled_frame_is_byte_identical(po, tmp_path, w, h):
    out = str(tmp_path / "tiled.jpg")
    mp.spawn(_worker, args=(2, _free_port(), w, h, 80, out), nprocs=2, join=True)
    from pixo_b200 import synthetic
    ref = po.jpeg_encode(synthetic.noise(w, h, 3, 42), w, h, po.RGB, 80, po.S420)
    assert open(out, "rb").read() == ref


def test_sharding_plans():
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
