"""Quick device-side timing of the transform kernels (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
import pixo_b200
from pixo_b200 import _lib, jpeg

lib = _lib.load()
ctx = pixo_b200.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
_, _, lq, cq = jpeg.quant_tables(80)


def run(w, h, n, ct=2, ss=1, reps=20):
    bpp = 3 if ct == 2 else 1
    px = torch.randint(0, 256, (n, h * w * bpp), dtype=torch.uint8, device="cuda")
    ny, nc = jpeg.block_counts(w, h, ct, ss)
    y = torch.empty((n, ny * 64), dtype=torch.int16, device="cuda")
    cb = torch.empty((n, max(nc, 1) * 64), dtype=torch.int16, device="cuda")
    cr = torch.empty_like(cb)
    def go():
        rc = lib.pixo_b200_jpeg_coefficients_dev(ctx.handle, px.data_ptr(), h * w * bpp, n, w, h, ct, ss,
                                                 lq.ctypes.data_as(_lib.f32p), cq.ctypes.data_as(_lib.f32p),
                                                 y.data_ptr(), ny * 64, cb.data_ptr(), cr.data_ptr(), max(nc, 1) * 64, 0, None)
        _lib.check(ctx.handle, rc)
    for _ in range(3): go()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byts = n * (h * w * bpp + (ny + 2 * nc) * 128)
    print(f"{w}x{h} x{n} ct={ct} ss={ss}: {ms*1e3:.1f} us/launch  {n*w*h/ms/1e3:.1f} Mpix/s  {byts/ms/1e6:.1f} GB/s")


if len(sys.argv) > 1 and sys.argv[1] == "short":
    run(3840, 2160, 1); run(3840, 2160, 32, reps=30); run(3840, 2160, 32, reps=30)
    sys.exit(0)
run(3840, 2160, 1)
run(3840, 2160, 16)
run(3840, 2160, 32)
run(1920, 1080, 64)
run(3840, 2160, 16, ss=0)
run(3840, 2160, 16, ct=0, ss=0)
run(1000, 1000, 16)
