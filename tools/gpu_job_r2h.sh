#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/quick_png.py 2>&1 | head -6
python tools/dbg_single.py 2>&1 | grep -E "call [3-4]" | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --configs C4,C5 > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2h.json'))
print('value',d['value'],'ms',d['ms_per_step'],'single',d['e2e']['single_call_pageable'])
for k,v in d['configs'].items(): print(k, json.dumps(v)[:1000])
PY
tail -5 gpurun_out/bench_r2h.err
