#!/bin/bash
# k_huff: software-pipelined symbol loop / branch-free window assembly, A/B + parity
mkdir -p gpurun_out
timeout 600 python tools/ab_huff.py pixo_b200/libpixo_b200.so tools/var/h_swp.so tools/var/h_sw2.so tools/var/h_swp_sw2.so pixo_b200/libpixo_b200.so 2>&1 | tee gpurun_out/ab_huff_r2p.txt
PIXO_B200_SO=$PWD/tools/var/h_swp_sw2.so timeout 1200 python -m pytest tests/test_jpeg_gpu.py tests/test_golden_reference.py tests/test_configs_full_gpu.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -4
