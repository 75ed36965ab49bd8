"""debug: fused Adler-32 of pixo_b200_png_filter_dev for various batch sizes / image sizes"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, zlib
import pixo_b200
from pixo_b200 import _lib
lib = _lib.load(); ctx = pixo_b200.Context(0)
for (w, h, n) in [(256, 64, 64), (256, 64, 129), (256, 64, 300), (3840, 2160, 100), (3840, 2160, 129), (3840, 2160, 132)]:
    rb = w * 4
    base = torch.randint(0, 256, (h * rb,), dtype=torch.uint8, device="cuda")
    d_in = base.repeat(n, 1).contiguous()
    d_out = torch.empty((n, h * (rb + 1)), dtype=torch.uint8, device="cuda")
    d_ad = torch.zeros(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    _lib.check(ctx.handle, lib.pixo_b200_png_filter_dev(ctx.handle, d_in.data_ptr(), h * rb, n, w, h, rb, 4, 6, d_out.data_ptr(), h * (rb + 1), d_ad.data_ptr()))
    ctx.sync()
    ad = d_ad.cpu().numpy().view(np.uint32)
    ref = zlib.adler32(d_out[0].cpu().numpy().tobytes())
    bad = np.nonzero(ad != ref)[0]
    print(w, h, n, "ref", hex(ref), "bad", len(bad), bad[:10], [hex(int(x)) for x in ad[bad[:4]]])
