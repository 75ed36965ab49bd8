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


def test_tiled_frame_entropy_coded_on_the_gpu(po, lib, gpu_ctx):
    """The bands' coefficients, gathered into device memory of one GPU, go through
    pixo_b200_jpeg_entropy_encode_dev (K3 + k_huff): same file as the oracle, also with a restart
    interval and optimised tables."""
    import torch
    import pixo_b200
    from pixo_b200 import ColorType, jpeg, parallel, synthetic
    from pixo_b200.jpeg import JpegOptions, Subsampling
    ndev = lib.pixo_b200_device_count()
    w, h, q = 2048, 1000, 80
    frame = synthetic.noise(w, h, 3, 42)
    bands = parallel.plan_bands(w, h, 8)
    ys, cbs, crs = [], [], []
    for b in bands:
        ctx = gpu_ctx if ndev == 1 else pixo_b200.Context(b.rank % ndev)
        px = np.ascontiguousarray(parallel.band_pixels(frame, w, h, 3, b)).reshape(-1)
        y, cb, cr = jpeg.compute_all_coefficients(px, w, b.px_row1 - b.px_row0, ColorType.Rgb, Subsampling.S420, q, ctx=ctx)
        ys.append(y); cbs.append(cb); crs.append(cr)
    dev = torch.device("cuda", gpu_ctx.device)
    d_y = torch.from_numpy(np.concatenate(ys)).to(dev)
    d_cb = torch.from_numpy(np.concatenate(cbs)).to(dev)
    d_cr = torch.from_numpy(np.concatenate(crs)).to(dev)
    torch.cuda.synchronize(dev)
    for ri, opt in ((None, False), (37, False), (None, True), (128, True)):
        o = JpegOptions(w, h, ColorType.Rgb, q, Subsampling.S420, ri, opt)
        got = jpeg.entropy_encode_dev(d_y, d_cb, d_cr, o, ctx=gpu_ctx)
        assert got == po.jpeg_encode(frame, w, h, po.RGB, q, po.S420, ri or 0, opt), (ri, opt)
