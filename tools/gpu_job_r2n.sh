#!/bin/bash
# packed lanes in k_huff (sparse images: 2 / 4 blocks per lane) + PNG emit clean-up: full parity suite, timings
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 300 python tools/huff_time.py 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2n.json 2> gpurun_out/bench_r2n.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2n.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'single',d['e2e']['single_call_pageable']['ms_per_call'],'launches',d['gpu_launches'],'roof',d['roofline']['frac'], d['roofline'].get('entropy_kernel_ms_per_step'))
for k,v in d['configs'].items(): print(k, json.dumps(v)[:900])
PY
tail -3 gpurun_out/bench_r2n.err
