#!/bin/bash
mkdir -p gpurun_out
python tools/dbg_single.py 2>&1 | tail -14
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2e.json'))
print('value',d['value'],'ms',d['ms_per_step'],'k1 frac',d['roofline']['frac'],'k1 ms',d['roofline']['kernel_ms_per_launch'],'huff ms',d['roofline']['entropy_kernel_ms_per_step'])
print('e2e',d['e2e']['value'],'single',d['e2e']['single_call_pageable'])
for k,v in d['configs'].items(): print(k, json.dumps(v)[:1000])
PY
tail -5 gpurun_out/bench_r2e.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_png_band -s 1 -c 1 -f -o gpurun_out/png_r2e python tools/prof_run.py png 16 3 > gpurun_out/ncu_png.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_huff -s 2 -c 1 -f -o gpurun_out/huff_r2e python tools/prof_run.py encode 32 4 > gpurun_out/ncu_huff.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_jpeg_420 -s 2 -c 1 -f -o gpurun_out/k1_r2e python tools/prof_run.py encode 32 4 > gpurun_out/ncu_k1.log 2>&1
tail -2 gpurun_out/ncu_png.log gpurun_out/ncu_huff.log gpurun_out/ncu_k1.log
