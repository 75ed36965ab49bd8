#!/bin/bash
# packed lanes: timing of the default (4/2/1), 2/1 and off builds, then parity for the default
mkdir -p gpurun_out
timeout 600 python tools/ab_huff.py pixo_b200/libpixo_b200.so tools/var/hp0.so tools/var/hp2.so pixo_b200/libpixo_b200.so 2>&1 | tee gpurun_out/ab_huff_r2o.txt
timeout 1200 python -m pytest tests/test_jpeg_gpu.py tests/test_golden_reference.py tests/test_configs_full_gpu.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -4
