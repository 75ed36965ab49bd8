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
print("tour done")
