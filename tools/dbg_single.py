"""debug: phases of one pageable single-frame encode through the C ABI (PIXO_B200_TIMING=1)"""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PIXO_B200_TIMING"] = "1"
import numpy as np
import pixo_b200
from pixo_b200 import _lib, synthetic
lib = _lib.load(); ctx = pixo_b200.Context(0)
w, h = 3840, 2160
for name, img in (("noise", synthetic.noise(w, h, 3, 43)), ("gradient", synthetic.gradient_rgb(w, h))):
    out = np.zeros(2 * img.size, np.uint8); n = C.c_size_t()
    for i in range(5):
        t = time.perf_counter()
        rc = lib.pixo_b200_jpeg_encode(ctx.handle, img.ctypes.data, img.size, w, h, 2, 80, 1, 0, 0, 0, 0, out.ctypes.data, out.size, C.byref(n))
        print(name, "call", i, round((time.perf_counter() - t) * 1e3, 3), "ms", n.value, rc, file=sys.stderr)
