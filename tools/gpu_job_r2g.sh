#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/dbg_single.py 2>&1 | grep -E "call [2-4]" | tail -6
python tools/huff_time.py 2>&1 | tail -2
PIXO_B200_SEGMENTS=1 python tools/huff_time.py 2>&1 | tail -1
PIXO_B200_SEGMENTS=2 python tools/huff_time.py 2>&1 | tail -1
PIXO_B200_SEGMENTS=4 python tools/huff_time.py 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 --configs C4 > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2g.json'))
print('value',d['value'],'ms',d['ms_per_step'],'single',d['e2e']['single_call_pageable'])
for k,v in d['configs'].items(): print(k, json.dumps(v)[:1000])
PY
tail -5 gpurun_out/bench_r2g.err
