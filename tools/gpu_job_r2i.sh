#!/bin/bash
mkdir -p gpurun_out
echo "=== k_huff occupancy variants ==="
timeout 600 python tools/ab_huff.py pixo_b200/libpixo_b200.so tools/var/libo7.so tools/var/libo6s12.so 2>&1 | tail -4
