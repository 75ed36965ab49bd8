"""Short launch sequences for ncu captures (never used for bench numbers)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pixo_b200
from pixo_b200 import _lib, jpeg

which = sys.argv[1] if len(sys.argv) > 1 else "jpeg420"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
lib = _lib.load()
ctx = pixo_b200.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
W, H = 3840, 2160
if which == "jpeg420":
    _, _, lq, cq = jpeg.quant_tables(80)
    px = torch.randint(0, 256, (n, H * W * 3), dtype=torch.uint8, device="cuda")
    ny, nc = jpeg.block_counts(W, H, 2, 1)
    y = torch.empty((n, ny * 64), dtype=torch.int16, device="cuda")
    cb = torch.empty((n, nc * 64), dtype=torch.int16, device="cuda"); cr = torch.empty_like(cb)
    for _ in range(reps):
        _lib.check(ctx.handle, lib.pixo_b200_jpeg_coefficients_dev(
            ctx.handle, px.data_ptr(), H * W * 3, n, W, H, 2, 1, lq.ctypes.data_as(_lib.f32p),
            cq.ctypes.data_as(_lib.f32p), y.data_ptr(), ny * 64, cb.data_ptr(), cr.data_ptr(), nc * 64, 0, None))
elif which == "encode":   # whole device path on the bench's frame mix (gradient / LCG noise)
    from pixo_b200 import synthetic
    g = synthetic.gradient_rgb(W, H).reshape(H, W * 3)
    fr = np.stack([np.roll(g, k, axis=0).reshape(-1) if k % 2 == 0 else synthetic.noise(W, H, 3, 42 + k).reshape(-1)
                   for k in range(n)])
    px = torch.from_numpy(fr).cuda()
    cap = (H * W * 3 // 2 + 65536 + 8192) // 256 * 256
    scan = torch.empty((n, cap), dtype=torch.uint8, device="cuda")
    sl = torch.zeros(n, dtype=torch.int64, device="cuda"); so = torch.zeros(n, dtype=torch.int32, device="cuda")
    for _ in range(reps):
        _lib.check(ctx.handle, lib.pixo_b200_jpeg_encode_dev(ctx.handle, px.data_ptr(), H * W * 3, n, W, H, 2, 80, 1,
                                                             scan.data_ptr(), cap, sl.data_ptr(), so.data_ptr()))
    torch.cuda.synchronize()
    print("scan bytes", sl.cpu().numpy().tolist(), "overflow", int(so.sum()))
elif which == "png":
    rb = W * 4
    px = torch.randint(0, 256, (n, H * rb), dtype=torch.uint8, device="cuda")
    out = torch.empty((n, H * (rb + 1)), dtype=torch.uint8, device="cuda")
    ad = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(reps):
        _lib.check(ctx.handle, lib.pixo_b200_png_filter_dev(ctx.handle, px.data_ptr(), H * rb, n, W, H, rb, 4, 6,
                                                            out.data_ptr(), H * (rb + 1), ad.data_ptr()))
torch.cuda.synchronize()
print("done", which, n, reps, ctx.launch_count)
