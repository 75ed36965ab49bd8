#!/bin/bash
# A/B of k_huff build variants (timing) + parity tests of the most complete one
mkdir -p gpurun_out
timeout 600 python tools/ab_huff.py tools/var/h_base.so tools/var/h_lb1.so tools/var/h_tick.so tools/var/h_pref.so tools/var/h_tick_pref.so tools/var/h_runs.so tools/var/h_runs_tp.so tools/var/h_base.so 2>&1 | tee gpurun_out/ab_huff_r2l.txt
PIXO_B200_SO=$PWD/tools/var/h_runs_tp.so timeout 900 python -m pytest tests/test_jpeg_gpu.py tests/test_golden_reference.py tests/test_configs_full_gpu.py -m gpu -x -q 2>&1 | tail -5
