"""One 16 384^2 frame through the single-context device path, a few times (for an ncu launch list:
which of K1 / k_huff<RAW> / k_seg_* the 1.6 ms go to).  Development aid, never a bench number."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pixo_b200
from pixo_b200 import _lib

lib = _lib.load()
ctx = pixo_b200.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
W = H = 16384
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
x = torch.arange(W, device="cuda", dtype=torch.int64)[None, :]
y = torch.arange(H, device="cuda", dtype=torch.int64)[:, None]
px = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
px[..., 0] = (x * 255 // W).to(torch.uint8)
px[..., 1] = (y * 255 // H).to(torch.uint8).expand(H, W)
px[..., 2] = ((x + y) * 127 // (W + H)).to(torch.uint8)
for s0 in (2040, 9000, 14336):   # three 2048-row noise stripes, like bench.py's C4 frame
    px[s0:s0 + 2048] = torch.randint(0, 256, (2048, W, 3), dtype=torch.uint8, device="cuda")
px = px.reshape(-1)
cap = (H * W * 3 // 2 + 65536 + 8192) // 256 * 256
scan = torch.empty(cap, dtype=torch.uint8, device="cuda")
sl = torch.zeros(1, dtype=torch.int64, device="cuda"); so = torch.zeros(1, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps):
    e0.record(stream)
    _lib.check(ctx.handle, lib.pixo_b200_jpeg_encode_dev(ctx.handle, px.data_ptr(), H * W * 3, 1, W, H, 2, 80, 1,
                                                         scan.data_ptr(), cap, sl.data_ptr(), so.data_ptr()))
    e1.record(stream); torch.cuda.synchronize()
    print(f"call {i}: {e0.elapsed_time(e1):.3f} ms, scan bytes {int(sl[0])}, flags {int(so[0])}")
