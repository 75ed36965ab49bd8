"""debug: phases of one pageable single-frame encode (PIXO_B200_TIMING=1)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PIXO_B200_TIMING"] = "1"
import numpy as np
import pixo_b200
from pixo_b200 import jpeg, synthetic, ColorType
from pixo_b200.jpeg import JpegOptions, Subsampling
ctx = pixo_b200.Context(0)
w, h = 3840, 2160
img = synthetic.noise(w, h, 3, 43)
o = JpegOptions(w, h, ColorType.Rgb, 80, Subsampling.S420)
for i in range(6):
    t = time.perf_counter(); b = jpeg.encode(img, o, ctx=ctx); print("call", i, round((time.perf_counter() - t) * 1e3, 3), "ms", len(b), file=sys.stderr)
