"""Device-side timings for SURVEY.md section 8's configurations C2-C5 (CUDA events on the launching stream,
3 warm-up launches, inputs resident in HBM and larger than L2).  Writes one JSON document:
    python tools/measure_configs.py > gpurun_out/configs.json
bench.py remains the contract measurement (C2); this covers the other configurations."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import pixo_b200
from pixo_b200 import _lib, jpeg, synthetic

lib = _lib.load()
ctx = pixo_b200.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
PEAK = 6568.4
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def frames(w, h, n, distinct=8):
    g = synthetic.gradient_rgb(w, h).reshape(h, w * 3)
    base = [np.roll(g, k, axis=0).reshape(-1) if k % 2 == 0 else synthetic.noise(w, h, 3, 42 + k) for k in range(min(n, distinct))]
    return torch.from_numpy(np.stack([base[k % len(base)] for k in range(n)])).cuda()


def jpeg_config(name, w, h, n, q, reps, ss=1):
    px = frames(w, h, n)
    nbytes = w * h * 3
    ny, nc = jpeg.block_counts(w, h, 2, ss)
    _, _, lq, cq = jpeg.quant_tables(q)
    y = torch.empty((n, ny * 64), dtype=torch.int16, device="cuda")
    cb = torch.empty((n, nc * 64), dtype=torch.int16, device="cuda"); cr = torch.empty_like(cb)
    cap = (nbytes * (1 if ss == 0 else 1) // (1 if ss == 0 else 2) + 65536 + 8192) // 256 * 256
    scan = torch.empty((n, cap), dtype=torch.uint8, device="cuda")
    sl = torch.zeros(n, dtype=torch.int64, device="cuda"); so = torch.zeros(n, dtype=torch.int32, device="cuda")

    def k1():
        _lib.check(ctx.handle, lib.pixo_b200_jpeg_coefficients_dev(
            ctx.handle, px.data_ptr(), nbytes, n, w, h, 2, ss, lq.ctypes.data_as(_lib.f32p), cq.ctypes.data_as(_lib.f32p),
            y.data_ptr(), ny * 64, cb.data_ptr(), cr.data_ptr(), nc * 64, 0, None))

    def enc():
        _lib.check(ctx.handle, lib.pixo_b200_jpeg_encode_dev(ctx.handle, px.data_ptr(), nbytes, n, w, h, 2, q, ss,
                                                             scan.data_ptr(), cap, sl.data_ptr(), so.data_ptr()))
    tk, te = timed(k1, reps), timed(enc, reps)
    assert int(so.sum()) == 0
    pix = n * w * h
    algo = n * (nbytes + (ny + 2 * nc) * 128)
    return {"config": name, "frames": n, "size": f"{w}x{h}", "quality": q, "subsampling": "4:2:0" if ss else "4:4:4",
            "k1_us": tk * 1e6, "k1_mpix_s": pix / tk / 1e6, "k1_gb_s": algo / tk / 1e9, "k1_frac_of_hbm_peak": algo / tk / 1e9 / PEAK,
            "device_path_us": te * 1e6, "device_path_mpix_s": pix / te / 1e6, "jpeg_bytes": int(sl.sum())}


def png_config(name, w, h, n, strat, reps):
    rb = w * 4
    g = synthetic.gradient_rgb(w, h).reshape(h, w, 3)
    rgba = np.concatenate([g, np.full((h, w, 1), 255, np.uint8)], -1).reshape(-1)
    base = [np.roll(rgba.reshape(h, rb), k, axis=0).reshape(-1) if k % 2 == 0 else synthetic.noise(w, h, 4, 42 + k) for k in range(4)]
    px = torch.from_numpy(np.stack([base[k % 4] for k in range(n)])).cuda()
    out = torch.empty((n, h * (rb + 1)), dtype=torch.uint8, device="cuda")
    ad = torch.empty(n, dtype=torch.int32, device="cuda")

    def go():
        _lib.check(ctx.handle, lib.pixo_b200_png_filter_dev(ctx.handle, px.data_ptr(), h * rb, n, w, h, rb, 4, strat,
                                                            out.data_ptr(), h * (rb + 1), ad.data_ptr()))
    t = timed(go, reps)
    algo = n * (h * rb + h * (rb + 1))
    return {"config": name, "images": n, "size": f"{w}x{h} RGBA", "strategy": strat, "us": t * 1e6,
            "mpix_s": n * w * h / t / 1e6, "gb_s": algo / t / 1e9, "frac_of_hbm_peak": algo / t / 1e9 / PEAK}


res = {"hbm_peak_gb_s": PEAK, "results": []}
res["results"].append(jpeg_config("C2 (32 x 3840x2160, q80)", 3840, 2160, 32, 80, 10))
res["results"].append(jpeg_config("pixo default preset (32 x 3840x2160, 4:4:4 q75)", 3840, 2160, 32, 75, 10, ss=0))
for q in (50, 80, 95):
    res["results"].append(jpeg_config(f"C3 (256 x 1920x1080, q{q})", 1920, 1080, 256, q, 5))
res["results"].append(jpeg_config("C4 (1 x 16384x16384, q80)", 16384, 16384, 1, 80, 5))
res["results"].append(png_config("C5 (64 x 3840x2160 RGBA, Adaptive)", 3840, 2160, 64, 6, 5))
res["results"].append(png_config("C5 (64 x 3840x2160 RGBA, AdaptiveFast)", 3840, 2160, 64, 7, 5))
print(json.dumps(res, indent=1))
