#!/bin/bash
mkdir -p gpurun_out
python tools/dbg_png.py 2>&1 | tail -8
echo "=== k_huff variants (emit, zrl-split) ==="
timeout 600 python tools/ab_huff.py tools/var/libh00.so tools/var/libh10.so tools/var/libh01.so tools/var/libh11.so 2>&1 | tail -8
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12
