"""Multi-GPU host logic: one process per GPU (torch.distributed), no collective on the data
path of batch encodes; one gather when a single huge frame is tiled across ranks.

  * batches (BASELINE configs C2/C3/C5): frame i -> rank i % world.  Ranks never exchange
    pixels or coefficients; only the timing reduction in bench.py is collective.
  * one gigapixel frame (config C4): contiguous bands of MCU rows per rank.  Every MCU depends
    only on its own (clamped) pixels, so a band is transformed as if it were an image of its
    own (bands start on multiples of 16/8 rows: no halo); the per-band coefficient arrays are
    gathered to rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests), concatenated in band
    order — that is exactly compute_all_coefficients' MCU order (src/jpeg/mod.rs:1046-1125) —
    and entropy-coded there.  The DC prediction chain crosses band boundaries inside the
    entropy stage, which sees the concatenated arrays, so the file is byte-identical to a
    single-GPU (and to the reference's) encode.
"""
from __future__ import annotations

import dataclasses

import numpy as np


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


def gather_coefficients(y, cb, cr, bands: list[Band], rank: int, world: int, dst: int = 0):
    """Gather per-band coefficient arrays (torch tensors on this rank's device) to `dst` and
    concatenate them in band order.  Returns (y, cb, cr) on dst, None elsewhere."""
    import torch
    import torch.distributed as dist

    def gather(t, counts):
        mx = max(counts) if counts else 0
        if world == 1:
            return t
        # collectives move raw bytes (gloo has no int16): view the coefficients as uint8
        pad = torch.zeros((mx, 128), dtype=torch.uint8, device=t.device)
        if t.shape[0]:
            pad[: t.shape[0]] = t.contiguous().view(torch.uint8).reshape(-1, 128)
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        if dist.get_backend() == "nccl":
            allb = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(allb, pad)
            bufs = allb if rank == dst else None
        else:
            dist.gather(pad, bufs, dst=dst)
        if rank != dst:
            return None
        return torch.cat([bufs[r][: counts[r]] for r in range(world)], 0).view(torch.int16).reshape(-1, 64)

    yc = [b.y_blocks for b in bands]
    cc = [b.c_blocks for b in bands]
    gy = gather(y, yc)
    gcb = gather(cb, cc) if any(cc) else None
    gcr = gather(cr, cc) if any(cc) else None
    return gy, gcb, gcr
