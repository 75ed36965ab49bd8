#!/bin/bash
mkdir -p gpurun_out
python tools/dbg_single.py 2>&1 | grep -E "call [2-4]|waited" | tail -12
python tools/quick_png.py 2>&1 | tail -10
timeout 900 python -m pytest tests/test_png_gpu.py tests/test_golden_reference.py tests/test_jpeg_gpu.py tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -5
