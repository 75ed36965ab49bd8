#!/bin/bash
# closing single-GPU run: full parity suite, smoke, the bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/bench_final.json') if l.startswith('{')][-1]
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'single',d['e2e']['single_call_pageable']['ms_per_call'],'launches',d['gpu_launches'],'roof',d['roofline']['frac'], d['roofline'].get('entropy_kernel_ms_per_step'), d['clocks'], 'fallbacks', d.get('host_fallbacks'))
for k,v in d['configs'].items(): print(k, json.dumps(v)[:500])
PY
tail -3 gpurun_out/bench_final.err
