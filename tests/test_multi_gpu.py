"""Config C4 on however many GPUs the box has (>= 1): a frame tiled into MCU-row bands, one
band per GPU context, transformed by the CUDA kernel, gathered in band order and entropy-coded —
byte-identical to the single-context encode and to the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_tiled_frame_bands_match_whole_frame(po, lib, gpu_ctx):
    import pixo_b200
    from pixo_b200 import ColorType, jpeg, parallel, synthetic
    from pixo_b200.jpeg import JpegOptions, Subsampling
    ndev = lib.pixo_b200_device_count()
    w, h, q = 2048, 1000, 80
    frame = synthetic.noise(w, h, 3, 42)
    world = 8
    bands = parallel.plan_bands(w, h, world)
    ys, cbs, crs = [], [], []
    for b in bands:
        ctx = gpu_ctx if ndev == 1 else pixo_b200.Context(b.rank % ndev)
        px = np.ascontiguousarray(parallel.band_pixels(frame, w, h, 3, b)).reshape(-1)
        y, cb, cr = jpeg.compute_all_coefficients(px, w, b.px_row1 - b.px_row0, ColorType.Rgb, Subsampling.S420, q, ctx=ctx)
        ys.append(y); cbs.append(cb); crs.append(cr)
    o = JpegOptions(w, h, ColorType.Rgb, q, Subsampling.S420)
    tiled = jpeg.entropy_encode(np.concatenate(ys), np.concatenate(cbs), np.concatenate(crs), o)
    assert tiled == jpeg.encode(frame, o, ctx=gpu_ctx) == po.jpeg_encode(frame, w, h, po.RGB, q, po.S420)
