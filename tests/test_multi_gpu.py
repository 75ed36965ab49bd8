"""Config C4 (one frame tiled into MCU-row bands) and C5 in row bands on the GPU: every band runs
the WHOLE path on its device (transform, K3, k_huff<RAW>, k_splice), only scan bytes are gathered.
With one device the bands run one after the other in one context; with >= 2 devices the same
frame also goes through torch.distributed/NCCL, one process per GPU."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _device_coders(ctx, frame, w, h, ct, ss, q, world):
    import torch
    from pixo_b200 import _lib, jpeg, parallel
    lib = _lib.load()
    ch = 1 if ct == 0 else 3
    dev = torch.device("cuda", ctx.device)
    _, _, lq, cq = jpeg.quant_tables(q)
    coders, keep = [], []
    for b in parallel.plan_bands(w, h, world, gray=ct == 0, s420=ss == 1):
        bh = b.px_row1 - b.px_row0
        ny, nc = b.y_blocks, b.c_blocks
        d_y = torch.empty((max(ny, 1), 64), dtype=torch.int16, device=dev)
        d_cb = torch.empty((max(nc, 1), 64), dtype=torch.int16, device=dev)
        d_cr = torch.empty((max(nc, 1), 64), dtype=torch.int16, device=dev)
        if bh:
            px = torch.from_numpy(np.ascontiguousarray(parallel.band_pixels(frame, w, h, ch, b)).reshape(-1)).to(dev)
            keep.append(px)
            torch.cuda.synchronize(dev)   # torch's stream is not the context's
            rc = lib.pixo_b200_jpeg_coefficients_dev(ctx.handle, px.data_ptr(), px.numel(), 1, w, bh, ct, ss,
                                                     lq.ctypes.data_as(_lib.f32p), cq.ctypes.data_as(_lib.f32p),
                                                     d_y.data_ptr(), ny * 64, d_cb.data_ptr(), d_cr.data_ptr(), nc * 64, 0, None)
            _lib.check(ctx.handle, rc)
        coders.append(parallel.DeviceBandCoder(ctx, d_y, d_cb if ct else None, d_cr if ct else None, w, max(bh, 1),
                                               ct, ss, ny, nc))
    ctx.sync()
    return coders, keep


@pytest.mark.parametrize("w,h,world,ct,ss", [(2048, 1000, 8, 2, 1), (333, 517, 5, 2, 0), (100, 40, 8, 2, 1),
                                             (1000, 777, 3, 0, 0)])
def test_tiled_frame_bands_on_one_gpu(po, gpu_ctx, w, h, world, ct, ss):
    from pixo_b200 import parallel, synthetic
    frame = synthetic.noise(w, h, 1 if ct == 0 else 3, 42)
    for q, opt in ((80, False), (100, False), (60, True)):
        coders, _keep = _device_coders(gpu_ctx, frame, w, h, ct, ss, q, world)
        got = parallel.encode_tiled_local(coders, w, h, ct, q, ss, opt)
        assert got == po.jpeg_encode(frame, w, h, ct, q, ss, 0, opt), (q, opt)
    assert gpu_ctx.host_fallbacks == 0


def test_tiled_frame_bands_coded_in_segments(po, gpu_ctx, monkeypatch):
    """Long bands are themselves cut into segments (a 16 384^2 frame on 8 GPUs: 16 per band); force
    it on a small frame."""
    from pixo_b200 import parallel, synthetic
    monkeypatch.setenv("PIXO_B200_SEGMENTS", "3")
    w, h = 640, 400
    frame = synthetic.noise(w, h, 3, 11)
    for q, opt in ((85, False), (100, True)):
        coders, _keep = _device_coders(gpu_ctx, frame, w, h, 2, 1, q, 4)
        assert parallel.encode_tiled_local(coders, w, h, 2, q, 1, opt) == po.jpeg_encode(frame, w, h, 2, q, 1, 0, opt)


@pytest.mark.parametrize("w,h,world,segs", [(1024, 640, 4, None), (640, 400, 3, "2"), (100, 40, 5, None)])
def test_stream_ordered_band_flow_with_thread_ranks(po, monkeypatch, w, h, world, segs):
    """The flow without host round trips (pixo_b200_jpeg_band_*_async: predictors, bit counts and
    offsets stay in device memory) with `world` ranks as threads of this process, a context and a
    stream each: the collectives are real exchanges between the ranks' device tensors."""
    import threading
    import torch
    import pixo_b200
    from pixo_b200 import parallel, synthetic
    if segs:
        monkeypatch.setenv("PIXO_B200_SEGMENTS", segs)
    frame = synthetic.noise(w, h, 3, 21)
    bands = parallel.plan_bands(w, h, world)
    nonempty = [b.y_blocks > 0 for b in bands]
    for q in (80, 100):
        comm = parallel.ThreadComm(world)
        res, errs = {}, []

        def work(rank):
            try:
                torch.cuda.set_device(0)
                stream = torch.cuda.Stream()
                with torch.cuda.stream(stream):
                    ctx = pixo_b200.Context(0)
                    ctx.set_stream(stream.cuda_stream)
                    coders, _keep = _device_coders(ctx, frame, w, h, 2, 1, q, world)
                    parts, _ = parallel.tiled_scan_parts_async(coders[rank], nonempty, rank, world, comm=comm)
                    if rank == 0:
                        res["jpg"] = parallel.assemble_tiled(parts, None, w, h, 2, q, 1)
                    assert ctx.host_fallbacks == 0
            except BaseException as e:   # noqa: BLE001 - re-raised in the main thread
                errs.append(e)
                comm.barrier.abort()

        ths = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        for t in ths: t.start()
        for t in ths: t.join()
        assert not errs, errs
        assert res["jpg"] == po.jpeg_encode(frame, w, h, 2, q, 1), q


def test_c4_full_size_16384_tiled_and_whole(po, gpu_ctx):
    """BASELINE config C4 at its stated size: one 16 384 x 16 384 RGB frame (805 MB, 6.3 M blocks),
    q=80 4:2:0 - eight bands with the distributed entropy stage AND the plain single-context encode,
    both byte-identical to the oracle."""
    import hashlib
    from pixo_b200 import ColorType, jpeg, parallel, synthetic
    from pixo_b200.jpeg import JpegOptions, Subsampling
    w = h = 16384
    g = synthetic.gradient_rgb(w, h).reshape(h, w * 3).copy()
    n = synthetic.noise(w, 2048, 3, 4242).reshape(2048, w * 3)
    for r0 in (2040, 9000, 14336):     # noise bands straddling / inside / ending a GPU band: dense + sparse content
        g[r0:r0 + 2048] = n
    frame = g.reshape(-1)
    ref = hashlib.sha256(po.jpeg_encode(frame, w, h, 2, 80, 1)).hexdigest()
    coders, _keep = _device_coders(gpu_ctx, frame, w, h, 2, 1, 80, 8)
    tiled = parallel.encode_tiled_local(coders, w, h, 2, 80, 1, False)
    assert hashlib.sha256(tiled).hexdigest() == ref
    del coders, _keep
    whole = jpeg.encode(frame, JpegOptions(w, h, ColorType.Rgb, 80, Subsampling.S420), ctx=gpu_ctx)
    assert hashlib.sha256(whole).hexdigest() == ref
    assert gpu_ctx.host_fallbacks == 0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _nccl_worker(rank, world, port, w, h, q, opt, out_path):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import pixo_b200
    from pixo_b200 import parallel, synthetic
    ctx = pixo_b200.Context(rank)
    frame = synthetic.noise(w, h, 3, 42)
    coders, _keep = _device_coders(ctx, frame, w, h, 2, 1, q, world)
    if opt:
        jpg = parallel.encode_tiled(coders[rank], w, h, 2, q, 1, opt, rank, world)
    else:   # the stream-ordered flow (needs the context on torch's current stream)
        stream = torch.cuda.Stream()
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            ctx.set_stream(stream.cuda_stream)
            bands = parallel.plan_bands(w, h, world)
            parts, _ = parallel.tiled_scan_parts_async(coders[rank], [b.y_blocks > 0 for b in bands], rank, world)
            jpg = parallel.assemble_tiled(parts, None, w, h, 2, q, 1) if rank == 0 else None
            stream.synchronize()
        ctx.set_stream(None)
    assert ctx.host_fallbacks == 0
    if rank == 0:
        open(out_path, "wb").write(jpg)
    dist.destroy_process_group()


def test_tiled_frame_over_nccl_on_real_devices(po, lib, tmp_path):
    """>= 2 GPUs: one process per GPU, band stages on its own device, collectives over NCCL."""
    import torch.multiprocessing as mp
    from pixo_b200 import synthetic
    ndev = lib.pixo_b200_device_count()
    if ndev < 2:
        pytest.skip("needs at least two CUDA devices")
    world = min(ndev, 8)
    w, h = 4096, 2048
    for q, opt in ((80, False), (90, True)):
        out = str(tmp_path / f"nccl_{q}.jpg")
        mp.spawn(_nccl_worker, args=(world, _free_port(), w, h, q, opt, out), nprocs=world, join=True)
        assert open(out, "rb").read() == po.jpeg_encode(synthetic.noise(w, h, 3, 42), w, h, 2, q, 1, 0, opt)


@pytest.mark.parametrize("strategy", ["Adaptive", "AdaptiveFast", "Paeth", "MinSum"])
def test_png_rows_in_bands_with_adler_combine(po, gpu_ctx, strategy):
    """SURVEY 8e PNG: one image cut into row bands, each band filtered with the raw row above it and
    check-summed on its own; the slices concatenate to the oracle's stream and the combined Adler-32
    equals the oracle's checksum of the whole stream."""
    import torch
    from pixo_b200 import parallel, png
    from pixo_b200.png import FilterStrategy
    w, h, bpp = 1000, 777, 4
    rb = w * bpp
    img = po.gen_noise(w, h, bpp, 9).reshape(h, rb).copy()
    img[100:300] = (np.arange(rb, dtype=np.uint32)[None, :] // 7 + np.arange(200, dtype=np.uint32)[:, None]).astype(np.uint8)
    strat = FilterStrategy[strategy]
    ref = po.apply_filters(img.reshape(-1), w, h, bpp, int(strat))
    dev = torch.device("cuda", gpu_ctx.device)
    d_img = torch.from_numpy(img).to(dev)
    cuts = [0, 1, 97, 400, 401, 776, h]
    outs, parts = [], []
    for r0, r1 in zip(cuts, cuts[1:]):
        rows = d_img[r0:r1].contiguous()
        above = d_img[r0 - 1].contiguous() if r0 else None
        d_out = torch.empty((r1 - r0) * (rb + 1), dtype=torch.uint8, device=dev)
        d_ad = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        png.apply_filters_rows_dev(rows, above, w, h, r1 - r0, rb, bpp, strat, d_out, d_ad, ctx=gpu_ctx)
        gpu_ctx.sync()
        outs.append(d_out.cpu().numpy())
        parts.append((int(d_ad.cpu().numpy().view(np.uint32)[0]), d_out.numel()))
    assert np.array_equal(np.concatenate(outs), ref)
    assert parallel.adler32_combine(parts) == po.adler32(ref)
