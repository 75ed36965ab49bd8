#!/bin/bash
# PNG emit clean-up: parity tests, timings, and the 3-CTA (80-register) variant
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_png_gpu.py tests/test_golden_reference.py tests/test_configs_full_gpu.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/quick_png.py 2>&1 | tail -20
timeout 300 python tools/ab_png.py tools/var/p_mb3.so pixo_b200/libpixo_b200.so 2>&1 | tail -4
timeout 600 python bench.py --configs C5 --steps 5 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
python - <<PY
import json
for l in open('gpurun_out/bench_c5.json'):
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps(d['configs']['C5']['strategies']))
PY
