import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PIXO_B200_SO"] = os.path.abspath("build/var/lib_ub.so")
import pixo_b200
from pixo_b200 import _lib
lib = _lib.load(); ctx = pixo_b200.Context(0)
lib.pixo_b200_ubench.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
iters = 200
for mode, name in ((0, "DCT+quant+stage only"), (1, "Y fill only"), (2, "Y fill + DCT+quant")):
    ms = C.c_float()
    rc = lib.pixo_b200_ubench(ctx.handle, mode, iters, C.byref(ms))
    # 148*3 CTAs * 4 warps, each `iters` block-jobs -> per SMSP: 3 warps * iters jobs
    cyc = ms.value * 1e-3 * 1.9e9 / (3 * iters)
    print(f"{name:24s} rc={rc} {ms.value:8.3f} ms  -> {cyc:7.0f} SMSP-cycles per warp-job (3 warps/SMSP sharing)")
