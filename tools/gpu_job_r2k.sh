#!/bin/bash
# two-phase PNG scoring: parity tests, then timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_png_gpu.py tests/test_golden_reference.py tests/test_configs_full_gpu.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python tools/quick_png.py 2>&1 | tail -20
timeout 600 python bench.py --configs C5 --steps 5 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
python - <<PY
import json
for l in open('gpurun_out/bench_c5.json'):
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps(d['configs'])[:1500])
PY
