#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
echo "=== K1 table-mode variants (huff_time: enc / k1 / huff us) ==="
timeout 600 python tools/ab_huff.py tools/var/libq0.so tools/var/libq1.so tools/var/libq2.so pixo_b200/libpixo_b200.so 2>&1 | tail -8
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2c.json'))
print('value',d['value'],'ms',d['ms_per_step'],'k1 frac',d['roofline']['frac'],'k1 ms',d['roofline']['kernel_ms_per_launch'],'huff ms',d['roofline']['entropy_kernel_ms_per_step'])
print('e2e',d['e2e']['value'],'single',d['e2e']['single_call_pageable'])
for k,v in d['configs'].items(): print(k, json.dumps(v)[:900])
PY
tail -5 gpurun_out/bench_r2c.err
