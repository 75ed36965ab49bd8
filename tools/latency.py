"""Host-call latency of pixo_b200_jpeg_encode on ordinary (pageable) input and output buffers,
through the C ABI (no Python-side copies)."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pixo_b200
from pixo_b200 import _lib, synthetic
lib = _lib.load(); ctx = pixo_b200.Context(0)
res = []
for (w, h, kind) in [(1920, 1080, "noise"), (3840, 2160, "noise"), (3840, 2160, "grad"), (8192, 8192, "grad")]:
    img = synthetic.gradient_rgb(w, h) if kind == "grad" else synthetic.noise(w, h, 3, 1)
    out = np.zeros(w * h * 2 + 4096, np.uint8)   # touched once: no first-touch faults in the timing
    n = C.c_size_t()
    def go():
        _lib.check(ctx.handle, lib.pixo_b200_jpeg_encode(ctx.handle, img.ctypes.data, img.size, w, h, 2, 80, 1, 0, 0, 0, 0,
                                                         out.ctypes.data, out.size, C.byref(n)))
    go(); go()
    t = time.perf_counter()
    for _ in range(8): go()
    dt = (time.perf_counter() - t) / 8
    res.append(f"{w}x{h} {kind}: {dt*1e3:.2f} ms ({w*h/dt/1e6:.0f} Mpix/s)")
print(" | ".join(res))
