"""A short tour of every kernel on small inputs, for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pixo_b200
from pixo_b200 import jpeg, png, synthetic, ColorType
from pixo_b200.jpeg import JpegOptions, Subsampling
from pixo_b200.png import FilterStrategy, PngOptions

ctx = pixo_b200.Context(0)
rng = np.random.default_rng(3)
for (w, h) in [(333, 222), (640, 480), (1000, 70)]:
    img = synthetic.noise(w, h, 3, 9)
    for ss in (Subsampling.S420, Subsampling.S444):
        for q, opt in ((80, False), (100, True)):
            out = jpeg.encode(img, JpegOptions(w, h, ColorType.Rgb, q, ss, None, opt), ctx=ctx)
            assert out[:2] == b"\xff\xd8" and out[-2:] == b"\xff\xd9"
    g = synthetic.noise(w, h, 1, 4)
    jpeg.encode(g, JpegOptions(w, h, ColorType.Gray, 85, Subsampling.S444), ctx=ctx)
    frames = np.stack([synthetic.noise(w, h, 3, k) for k in range(5)])
    jpeg.encode_batch(frames, JpegOptions(w, h, ColorType.Rgb, 80, Subsampling.S420), ctx=ctx)
    for bpp, ct in ((4, 3), (2, 1), (3, 2)):
        px = rng.integers(0, 256, (h, w, bpp), dtype=np.uint8)
        if bpp in (2, 4):
            px[..., -1] = np.where(rng.random((h, w)) < 0.3, 0, px[..., -1])
        for st in (FilterStrategy.Adaptive, FilterStrategy.AdaptiveFast, FilterStrategy.Bigrams, FilterStrategy.Paeth):
            png.apply_filters(px.reshape(-1), w, h, bpp, PngOptions(w, h, ColorType(ct), st, True), with_adler=True, ctx=ctx)
    png.adler32(img, ctx=ctx)
# ---- round 2: restart intervals, segmented coding (k_huff<RAW> + k_seg_*), bands of one frame (band entropy +
# splice, also segmented), the PNG row-band entry point and the mixed-row scoring paths ----------------------
import torch
from pixo_b200 import _lib, parallel
lib = _lib.load()
w, h = 1024, 512
img = synthetic.noise(w, h, 3, 11)
img.reshape(h, w * 3)[100:300] = synthetic.gradient_rgb(w, h).reshape(h, w * 3)[100:300]
plain = jpeg.encode(img, JpegOptions(w, h, ColorType.Rgb, 80, Subsampling.S420), ctx=ctx)
jpeg.encode(img, JpegOptions(w, h, ColorType.Rgb, 80, Subsampling.S420, 7), ctx=ctx)          # restart interval
for S in ("3", "16", "200"):
    os.environ["PIXO_B200_SEGMENTS"] = S
    seg = jpeg.encode(img, JpegOptions(w, h, ColorType.Rgb, 80, Subsampling.S420), ctx=ctx)
    assert seg == plain, S
    # bands of the frame, each band's raw string itself segmented
    dev = torch.device("cuda", 0)
    _, _, lq, cq = jpeg.quant_tables(80)
    coders, keep = [], []
    for b in parallel.plan_bands(w, h, 3):
        bh = b.px_row1 - b.px_row0
        d_y = torch.empty((max(b.y_blocks, 1), 64), dtype=torch.int16, device=dev)
        d_cb = torch.empty((max(b.c_blocks, 1), 64), dtype=torch.int16, device=dev)
        d_cr = torch.empty_like(d_cb)
        px = torch.from_numpy(np.ascontiguousarray(parallel.band_pixels(img, w, h, 3, b)).reshape(-1)).to(dev)
        keep.append(px)
        torch.cuda.synchronize(dev)
        _lib.check(ctx.handle, lib.pixo_b200_jpeg_coefficients_dev(
            ctx.handle, px.data_ptr(), px.numel(), 1, w, bh, 2, 1, lq.ctypes.data_as(_lib.f32p), cq.ctypes.data_as(_lib.f32p),
            d_y.data_ptr(), b.y_blocks * 64, d_cb.data_ptr(), d_cr.data_ptr(), b.c_blocks * 64, 0, None))
        coders.append(parallel.DeviceBandCoder(ctx, d_y, d_cb, d_cr, w, bh, 2, 1, b.y_blocks, b.c_blocks))
    ctx.sync()
    tiled = parallel.encode_tiled_local(coders, w, h, 2, 80, 1)
    assert tiled == plain, ("bands", S)
del os.environ["PIXO_B200_SEGMENTS"]
# PNG: flat rows next to noise rows (both scoring routes and their switches), and a row band with the row above
hh, ww, bpp = 70, 1024, 4
rows = rng.integers(0, 256, (hh, ww * bpp), dtype=np.uint8)
rows[::3] = 17
for st in (FilterStrategy.Adaptive, FilterStrategy.AdaptiveFast, FilterStrategy.MinSum):
    whole, _ = png.apply_filters(rows.reshape(-1), ww, hh, bpp, PngOptions(ww, hh, ColorType.Rgba, st), with_adler=True, ctx=ctx)
    d_rows = torch.from_numpy(rows).to(dev)
    torch.cuda.synchronize(dev)
    d_out = torch.empty(30 * (ww * bpp + 1), dtype=torch.uint8, device=dev)
    d_ad = torch.zeros(1, dtype=torch.int32, device=dev)
    torch.cuda.synchronize(dev)
    png.apply_filters_rows_dev(d_rows[20:50].contiguous().reshape(-1), d_rows[19].contiguous(), ww, hh, 30, ww * bpp, bpp, st,
                               d_out, d_ad, ctx=ctx)
    ctx.sync()
    assert np.array_equal(d_out.cpu().numpy(), whole.reshape(hh, ww * bpp + 1)[20:50].reshape(-1)), st
print("tour done")
