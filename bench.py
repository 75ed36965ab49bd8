#!/usr/bin/env python
"""bench.py — measures BASELINE.json's metric: Mpixels/s of JPEG q=80 4:2:0 encode of
3840x2160 RGB frames (config C2), per GPU and aggregated over N GPUs (one process per GPU).

A "step" is one pass of the hot path over one batch of FRAMES distinct synthetic 4K frames
(a ring larger than the 126 MB L2, so no step finds its input or output resident).

 value     device-resident throughput of the whole hot path: pixo_b200_jpeg_encode_dev =
           fused colour->subsample->DCT->quantise kernel + single-pass Huffman/stuffing
           kernel, RGB frames already in HBM -> entropy-coded scan bytes left in HBM; Mpix/s.
 e2e       the same metric through the reference-facing C ABI call with HOST buffers:
           pixo_b200_jpeg_encode_batch(pinned RGB frames) -> finished JPEG byte streams in host
           memory; H2D of the pixels, transform + Huffman/stuffing kernels, D2H of the scan bytes
           and the host-side header writing are all inside the timed region.
 roofline  achieved HBM GB/s of the transform kernel (algorithmic 6 B/px) vs MEASURED_PEAKS.
 cpu_baseline  the CPU restatement of pixo's encoder (oracle/, kind "port" — no Rust toolchain
           exists to build pixo itself) on the host cores, bounded sample.

`--impl reference` times that CPU encoder alone (all host threads, one frame per thread).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, QUALITY = 3840, 2160, 80
PIX = W * H
IN_BYTES = PIX * 3
ALGO_BYTES_PER_FRAME = IN_BYTES + (4 + 1 + 1) * (W // 16) * (H // 16) * 128  # 6 B/px = 49 766 400
METRIC = "Mpixels/sec JPEG q=80 4:2:0 encode of 3840x2160 RGB"


def make_frames(n: int) -> np.ndarray:
    """Ring of n distinct frames: even k = gradient shifted by k rows, odd k = LCG noise seed 42+k
    (SURVEY.md §8d generators)."""
    from pixo_b200 import synthetic
    g = synthetic.gradient_rgb(W, H).reshape(H, W * 3)
    out = np.empty((n, IN_BYTES), np.uint8)
    for k in range(n):
        if k % 2 == 0:
            out[k] = np.roll(g, k, axis=0).reshape(-1)
        else:
            out[k] = synthetic.noise(W, H, 3, 42 + k)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        mhz, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                mhz.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(mhz)}


def measured_peak_gbs() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------
def cpu_reference_rate(frames: np.ndarray, threads: int, rounds: int) -> tuple[float, float, int]:
    """The CPU encoder (oracle restatement of pixo::jpeg::encode, q80 4:2:0 baseline) on
    `threads` host threads, one whole frame per thread per round (pixo's baseline encode_scan is
    single-threaded per image, so frame-level parallelism is all the host threads it can use).
    Returns (Mpix/s, seconds, frames encoded)."""
    from oracle import pyoracle as po
    po.lib()
    n = frames.shape[0]
    bufs = [np.empty(po._jpeg_cap(W, H) // 16, np.uint8) for _ in range(threads)]
    done = [0] * threads

    def work(t):
        for r in range(rounds):
            k = (t + r * threads) % n
            cap = bufs[t].size
            ln = po.lib().po_jpeg_encode(frames[k].ctypes.data_as(po.u8p), IN_BYTES, W, H, po.RGB, QUALITY,
                                         po.S420, 0, 0, bufs[t].ctypes.data_as(po.u8p), cap)
            assert ln > 0, ln
            done[t] += 1

    ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for th in ths: th.start()
    for th in ths: th.join()
    dt = time.perf_counter() - t0
    nf = sum(done)
    return nf * PIX / dt / 1e6, dt, nf


def host_threads() -> int:
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0)) or n
    except AttributeError:
        pass
    return n


def best_thread_count(frames: np.ndarray) -> tuple[int, dict]:
    """The CPU arm gets the thread count it runs fastest with: all hardware threads, or one per
    physical core when SMT siblings only fight over the FP units (one calibration round each)."""
    n = host_threads()
    tried = {}
    for t in sorted({n, max(1, n // 2)}, reverse=True):
        tried[t] = cpu_reference_rate(frames, t, 1)[0]
    return max(tried, key=tried.get), tried


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = make_frames(min(16, max(2, host_threads())))
    threads, tried = best_thread_count(frames)   # doubles as the warm-up
    t_total, f_total = 0.0, 0
    for _ in range(args.steps):
        _, dt, nf = cpu_reference_rate(frames, threads, 1)
        t_total += dt; f_total += nf
    rate = f_total * PIX / t_total / 1e6
    sample = (f"{threads} frames of 3840x2160 per step (one per thread), {args.steps} steps; thread count chosen "
              f"by calibration {({k: round(v) for k, v in tried.items()})} Mpix/s")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: 3840x2160 RGB -> JPEG q=80 4:2:0 baseline, standard Huffman "
                               "(frames: gradient shifted / LCG noise alternating)",
                   "frames_per_step": threads, "note": "CPU restatement of pixo's encoder "
                   "(oracle/pixo_oracle.c, gcc -O2 strict binary32); pixo itself cannot be built here (no Rust)"},
        "cpu_baseline": {"value": rate, "unit": "Mpix/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    import pixo_b200
    from pixo_b200 import _lib, jpeg
    lib = _lib.load()
    ctx = pixo_b200.Context(local_rank)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    F = args.frames
    frames_host = make_frames(F)
    pinned = torch.from_numpy(frames_host).pin_memory()
    d_px = pinned.to(dev, non_blocking=True)
    ny, nc = jpeg.block_counts(W, H, 2, 1)
    d_y = torch.empty((F, ny * 64), dtype=torch.int16, device=dev)
    d_cb = torch.empty((F, nc * 64), dtype=torch.int16, device=dev)
    d_cr = torch.empty((F, nc * 64), dtype=torch.int16, device=dev)
    _, _, lq, cq = jpeg.quant_tables(QUALITY)
    lqp, cqp = lq.ctypes.data_as(_lib.f32p), cq.ctypes.data_as(_lib.f32p)

    scan_cap = IN_BYTES // 2 + 65536 + 8192
    scan_cap -= scan_cap % 256
    d_scan = torch.empty((F, scan_cap), dtype=torch.uint8, device=dev)
    d_slen = torch.zeros(F, dtype=torch.int64, device=dev)
    d_sovf = torch.zeros(F, dtype=torch.int32, device=dev)

    def encode_step():
        rc = lib.pixo_b200_jpeg_encode_dev(ctx.handle, d_px.data_ptr(), IN_BYTES, F, W, H, 2, QUALITY, 1,
                                           d_scan.data_ptr(), scan_cap, d_slen.data_ptr(), d_sovf.data_ptr())
        _lib.check(ctx.handle, rc)

    def kernel_step():
        rc = lib.pixo_b200_jpeg_coefficients_dev(ctx.handle, d_px.data_ptr(), IN_BYTES, F, W, H, 2, 1, lqp, cqp,
                                                 d_y.data_ptr(), ny * 64, d_cb.data_ptr(), d_cr.data_ptr(),
                                                 nc * 64, 0, None)   # natural order: the launch an encode step makes
        _lib.check(ctx.handle, rc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident whole path (value) ----
    for _ in range(max(args.warmup, 3)):
        encode_step()
    barrier()
    assert int(d_sovf.sum()) == 0, "scan capacity overflow"
    launches_enc0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        encode_step()
    ev1.record(stream)
    barrier()
    enc_launches = ctx.launch_count - launches_enc0
    tenc = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tenc, op=dist.ReduceOp.MAX)
    enc_ms = float(tenc[0])

    # ---- the dominant kernel alone (roofline) ----
    for _ in range(3):
        kernel_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e_all0.record(stream)
    for a, b in evs:
        a.record(stream); kernel_step(); b.record(stream)
    e_all1.record(stream)
    barrier()
    kernel_launches = ctx.launch_count - launches0
    total_ms = e_all0.elapsed_time(e_all1)
    per_launch_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    t = torch.tensor([total_ms, per_launch_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, per_launch_ms = float(t[0]), float(t[1])
    value = world * F * args.steps * PIX / (enc_ms * 1e-3) / 1e6
    k1_value = world * F * args.steps * PIX / (total_ms * 1e-3) / 1e6

    # ---- end-to-end through the C ABI with host buffers ----
    e2e_frames = min(F, args.e2e_frames)
    cap = 2 * IN_BYTES
    out_host = torch.empty((e2e_frames, cap), dtype=torch.uint8).pin_memory()
    lens = (C.c_size_t * e2e_frames)()

    def e2e_step():
        rc = lib.pixo_b200_jpeg_encode_batch(ctx.handle, pinned.data_ptr(), IN_BYTES, e2e_frames, W, H, 2, QUALITY,
                                             1, 0, 0, out_host.data_ptr(), cap, lens)
        _lib.check(ctx.handle, rc)

    e2e_step()
    barrier()
    launches1 = ctx.launch_count
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_launches = ctx.launch_count - launches1
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_frames * e2e_steps * PIX / float(te[0]) / 1e6
    jpeg_bytes = int(sum(lens))
    clocks = sampler.stop() if rank == 0 else None

    # ---- CPU baseline, rank 0 at N=1 only: bounded sample ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores, _tried = best_thread_count(frames_host)
        rate, dt, nf = cpu_reference_rate(frames_host, cores, 1)
        cpu = {"value": rate, "unit": "Mpix/s", "cores": cores, "kind": "port",
               "sample": f"{nf} frames of 3840x2160 (one per thread), {dt:.1f} s wall; CPU restatement of "
                         "pixo::jpeg::encode (oracle/), pixo itself unbuildable here (no Rust toolchain)"}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        achieved = ALGO_BYTES_PER_FRAME * F / (per_launch_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("k_jpeg_420_dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": enc_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C2: 3840x2160 RGB -> JPEG q=80 4:2:0; step = ring of {F} distinct frames per GPU "
                                   f"({F * IN_BYTES / 1e6:.0f} MB in + same out, larger than L2: no flush needed)",
                       "frames_per_step_per_gpu": F, "global_frames_per_step": world * F,
                       "frame_content": "even: gradient_rgb shifted k rows; odd: LCG noise seed 42+k",
                       "l2_policy": "inputs+outputs larger than L2", "parallelism": f"dp{world} (frames sharded, no collective on the data path)"},
            "roofline": {"bound": "hbm", "kernel": "k_jpeg_420", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_FRAME * F,
                         "kernel_ms_per_launch": per_launch_ms, "kernel_only_mpix_s": k1_value,
                         "share_of_step": per_launch_ms / (enc_ms / args.steps),
                         "entropy_kernel_ms_per_step": enc_ms / args.steps - per_launch_ms,
                         "note": "k_jpeg_420 timed alone (pixo_b200_jpeg_coefficients_dev, the same launch an encode "
                                 "step makes) on the same ring; the rest of a step is k_huff, the single-pass "
                                 "Huffman/stuffing kernel (instruction-issue bound, not HBM bound)"},
            "e2e": {"value": e2e_value, "unit": "Mpix/s", "h2d_bytes_per_step": e2e_frames * IN_BYTES,
                    "d2h_bytes_per_step": jpeg_bytes + e2e_frames * 12,
                    "frames_per_step": e2e_frames, "steps": e2e_steps, "jpeg_bytes_last_step": jpeg_bytes,
                    "api": "pixo_b200_jpeg_encode_batch (host RGB in pinned memory -> JPEG bytes on host)"},
            "gpu_launches": int(enc_launches + kernel_launches + e2e_launches),
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=32, help="distinct 4K frames per step per GPU")
    ap.add_argument("--e2e-frames", type=int, default=32)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
