"""BASELINE configs C3 and C5 at their STATED batch sizes (1.59 GB / 2.12 GB of input, i.e. byte
offsets beyond 2^31 inside one call), checked against the oracle on the first, a middle and the last
frame, plus pixo's default preset (4:4:4 q75) on 4K."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_c3_full_batch_256x1080p(po, gpu_ctx):
    from pixo_b200 import ColorType, jpeg, synthetic
    from pixo_b200.jpeg import JpegOptions, Subsampling
    w, h, n = 1920, 1080, 256
    g = synthetic.gradient_rgb(w, h).reshape(h, w * 3)
    frames = np.empty((n, w * h * 3), np.uint8)
    for k in range(n):
        frames[k] = synthetic.noise(w, h, 3, 42 + k) if k % 2 else np.roll(g, k, axis=0).reshape(-1)
    assert frames.nbytes > 1 << 30
    for q in (50, 80, 95):
        got = jpeg.encode_batch(frames, JpegOptions(w, h, ColorType.Rgb, q, Subsampling.S420), ctx=gpu_ctx)
        for k in (0, 1, 127, 254, 255):
            assert got[k] == po.jpeg_encode(frames[k], w, h, 2, q, 1), (q, k)
    assert gpu_ctx.host_fallbacks == 0


def test_c5_full_batch_64x4k_rgba(po, gpu_ctx):
    import torch
    from pixo_b200 import _lib, synthetic
    lib = _lib.load()
    w, h, n, bpp = 3840, 2160, 64, 4
    rb = w * bpp
    dev = torch.device("cuda", gpu_ctx.device)
    base = synthetic.gradient_rgb(w, h).reshape(h, w, 3)
    d_in = torch.empty((n, h * rb), dtype=torch.uint8, device=dev)
    hosts = {}
    for k in range(n):
        if k % 2:
            f = synthetic.noise(w, h, 4, 42 + k)
        else:
            f = np.concatenate([np.roll(base, k, axis=0), np.full((h, w, 1), 255, np.uint8)], -1).reshape(-1)
        if k in (0, 1, 31, 62, 63):
            hosts[k] = f
        d_in[k] = torch.from_numpy(f).to(dev)
    assert d_in.numel() > 1 << 30
    out_stride = h * (rb + 1)
    d_out = torch.empty((n, out_stride), dtype=torch.uint8, device=dev)
    d_ad = torch.zeros(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()   # torch fills ran on torch's stream; the library has its own
    for strat, code in (("adaptive", po.F_ADAPTIVE), ("fast", po.F_ADAPTIVE_FAST)):
        rc = lib.pixo_b200_png_filter_dev(gpu_ctx.handle, d_in.data_ptr(), h * rb, n, w, h, rb, bpp, code,
                                          d_out.data_ptr(), out_stride, d_ad.data_ptr())
        _lib.check(gpu_ctx.handle, rc)
        gpu_ctx.sync()
        ad = d_ad.cpu().numpy().view(np.uint32)
        for k, f in hosts.items():
            ref = po.apply_filters(f, w, h, bpp, code)
            assert hashlib.sha256(d_out[k].cpu().numpy().tobytes()).digest() == hashlib.sha256(ref.tobytes()).digest(), (strat, k)
            assert int(ad[k]) == po.adler32(ref), (strat, k)


def test_default_preset_444_q75_on_4k(po, gpu_ctx):
    """pixo's default / `fast` preset is 4:4:4 q75 (src/jpeg/mod.rs:142-174)."""
    from pixo_b200 import jpeg, synthetic
    from pixo_b200.jpeg import JpegOptions
    w, h = 3840, 2160
    for img in (synthetic.noise(w, h, 3, 3), synthetic.gradient_rgb(w, h)):
        o = JpegOptions.fast(w, h, 75)
        assert jpeg.encode(img, o, ctx=gpu_ctx) == po.jpeg_encode(img, w, h, 2, 75, 0)


def test_offsets_beyond_4_gib(po, gpu_ctx):
    """One call over more than 2^32 bytes of input AND output: 180 4K RGB frames through the JPEG
    device path (4.48 GB in) and 132 4K RGBA frames through the PNG filter (4.38 GB in/out); the last
    frames - whose byte offsets do not fit 32 bits - must match the oracle."""
    import torch
    from pixo_b200 import _lib, synthetic
    lib = _lib.load()
    dev = torch.device("cuda", gpu_ctx.device)
    w, h = 3840, 2160
    base = [synthetic.noise(w, h, 3, 5), synthetic.gradient_rgb(w, h)]
    n = 180
    d_px = torch.empty((n, w * h * 3), dtype=torch.uint8, device=dev)
    d_base = [torch.from_numpy(b).to(dev).reshape(h, -1) for b in base]
    for k in range(n):
        d_px[k] = torch.roll(d_base[k % 2], k, 0).reshape(-1)
    assert d_px.numel() > 1 << 32
    cap = (w * h * 3 // 2 + 65536) // 256 * 256
    d_scan = torch.empty((n, cap), dtype=torch.uint8, device=dev)
    d_len = torch.zeros(n, dtype=torch.int64, device=dev)
    d_ovf = torch.zeros(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()   # torch fills ran on torch's stream; the library has its own
    _lib.check(gpu_ctx.handle, lib.pixo_b200_jpeg_encode_dev(gpu_ctx.handle, d_px.data_ptr(), w * h * 3, n, w, h, 2, 80, 1,
                                                             d_scan.data_ptr(), cap, d_len.data_ptr(), d_ovf.data_ptr()))
    gpu_ctx.sync()
    assert int(d_ovf.sum()) == 0
    lens = d_len.cpu().numpy()
    for k in (0, 178, 179):
        ref = po.jpeg_encode(np.roll(base[k % 2].reshape(h, -1), k, axis=0).reshape(-1), w, h, 2, 80, 1)
        got = d_scan[k, : int(lens[k])].cpu().numpy().tobytes()
        assert got == ref[ref.index(b"\xff\xda") + 14:-2], k
    del d_px, d_scan
    torch.cuda.empty_cache()
    n, bpp = 132, 4
    rb = w * bpp
    rgba = [synthetic.noise(w, h, 4, 8), np.concatenate([synthetic.gradient_rgb(w, h).reshape(h, w, 3),
                                                         np.full((h, w, 1), 200, np.uint8)], -1).reshape(-1)]
    d_base = [torch.from_numpy(b).to(dev).reshape(h, rb) for b in rgba]
    d_in = torch.empty((n, h * rb), dtype=torch.uint8, device=dev)
    for k in range(n):
        d_in[k] = torch.roll(d_base[k % 2], k, 0).reshape(-1)
    assert d_in.numel() > 1 << 32
    d_out = torch.empty((n, h * (rb + 1)), dtype=torch.uint8, device=dev)
    d_ad = torch.zeros(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    _lib.check(gpu_ctx.handle, lib.pixo_b200_png_filter_dev(gpu_ctx.handle, d_in.data_ptr(), h * rb, n, w, h, rb, bpp,
                                                            po.F_ADAPTIVE, d_out.data_ptr(), h * (rb + 1), d_ad.data_ptr()))
    gpu_ctx.sync()
    ad = d_ad.cpu().numpy().view(np.uint32)
    for k in (0, 130, 131):
        ref = po.apply_filters(np.roll(rgba[k % 2].reshape(h, rb), k, axis=0).reshape(-1), w, h, bpp, po.F_ADAPTIVE)
        assert hashlib.sha256(d_out[k].cpu().numpy().tobytes()).digest() == hashlib.sha256(ref.tobytes()).digest(), k
        assert int(ad[k]) == po.adler32(ref), k
