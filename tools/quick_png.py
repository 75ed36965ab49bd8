"""Quick device-side timing of the PNG filter + Adler kernels (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pixo_b200
from pixo_b200 import _lib
lib = _lib.load()
ctx = pixo_b200.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)

def run(w, h, bpp, n, strat, smooth=False, reps=10):
    rb = w * bpp
    if smooth:
        base = torch.randint(0, 4, (n, h, rb), dtype=torch.int32, device="cuda").cumsum(2)
        px = (base & 255).to(torch.uint8).reshape(n, -1).contiguous()
    else:
        px = torch.randint(0, 256, (n, h * rb), dtype=torch.uint8, device="cuda")
    out = torch.empty((n, h * (rb + 1)), dtype=torch.uint8, device="cuda")
    ad = torch.empty(n, dtype=torch.int32, device="cuda")
    def go():
        _lib.check(ctx.handle, lib.pixo_b200_png_filter_dev(ctx.handle, px.data_ptr(), h * rb, n, w, h, rb, bpp, strat,
                                                            out.data_ptr(), h * (rb + 1), ad.data_ptr()))
    for _ in range(3): go()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byts = n * (h * rb + h * (rb + 1))
    print(f"png {w}x{h} bpp{bpp} x{n} strat={strat} {'smooth' if smooth else 'noise'}: {ms*1e3:.1f} us  {n*w*h/ms/1e3:.1f} Mpix/s  {byts/ms/1e6:.1f} GB/s")

def adler(nbytes, reps=10):
    d = torch.randint(0, 256, (nbytes,), dtype=torch.uint8, device="cuda")
    o = torch.empty(1, dtype=torch.int32, device="cuda")
    def go(): _lib.check(ctx.handle, lib.pixo_b200_adler32_dev(ctx.handle, d.data_ptr(), nbytes, o.data_ptr()))
    for _ in range(3): go()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"adler32 {nbytes} B: {ms*1e3:.1f} us  {nbytes/ms/1e6:.1f} GB/s")

N = int(os.environ.get("PNG_N", "16"))   # frames per call
run(3840, 2160, 4, N, 6)
run(3840, 2160, 4, N, 6, smooth=True)
run(3840, 2160, 4, N, 7)
run(3840, 2160, 4, 16, 4)
run(3840, 2160, 4, 16, 1)
run(3840, 2160, 3, 16, 6)
run(1000, 1000, 3, 16, 6)
adler(33179760)
adler(1 << 30)
