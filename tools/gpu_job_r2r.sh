#!/bin/bash
# segmented coding with the fast splice: does it pay for 4K batches now?  (test hook PIXO_B200_SEGMENTS)
mkdir -p gpurun_out
for S in 1 2 4 8 16; do echo "== segments per image: $S"; PIXO_B200_SEGMENTS=$S timeout 300 python tools/huff_time.py 2>&1 | tail -1 | cut -c1-260; done | tee gpurun_out/seg_sweep_r2r.txt
for S in 16 32 64; do echo "== 16K frame, segments: $S"; PIXO_B200_SEGMENTS=$S timeout 300 python tools/prof_c4.py 4 2>&1 | tail -2; done | tee -a gpurun_out/seg_sweep_r2r.txt
