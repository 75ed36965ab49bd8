#!/bin/bash
# tests, bench (with configs), ncu captures of the two JPEG kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err
tail -c 6000 gpurun_out/bench_r2b.json; tail -20 gpurun_out/bench_r2b.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_jpeg_420 -s 2 -c 1 -f -o gpurun_out/k1_r2b python tools/prof_run.py encode 32 4 > gpurun_out/ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_huff -s 2 -c 1 -f -o gpurun_out/huff_r2b python tools/prof_run.py encode 32 4 > gpurun_out/ncu_huff.log 2>&1
tail -3 gpurun_out/ncu_k1.log gpurun_out/ncu_huff.log
