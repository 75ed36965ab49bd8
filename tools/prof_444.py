import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pixo_b200
from pixo_b200 import _lib, jpeg
lib = _lib.load(); ctx = pixo_b200.Context(0)
W, H, n = 3840, 2160, 16
_, _, lq, cq = jpeg.quant_tables(80)
px = torch.randint(0, 256, (n, H * W * 3), dtype=torch.uint8, device="cuda")
ny, nc = jpeg.block_counts(W, H, 2, 0)
y = torch.empty((n, ny * 64), dtype=torch.int16, device="cuda"); cb = torch.empty_like(y); cr = torch.empty_like(y)
for _ in range(3):
    _lib.check(ctx.handle, lib.pixo_b200_jpeg_coefficients_dev(ctx.handle, px.data_ptr(), H * W * 3, n, W, H, 2, 0,
               lq.ctypes.data_as(_lib.f32p), cq.ctypes.data_as(_lib.f32p), y.data_ptr(), ny * 64, cb.data_ptr(), cr.data_ptr(), nc * 64, 0, None))
torch.cuda.synchronize(); print("done")
