#!/bin/bash
# 256 segments + parallel segment prefix: parity (segments, bands, C4 full size) and timing
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_jpeg_gpu.py tests/test_multi_gpu.py tests/test_configs_full_gpu.py tests/test_golden_reference.py -m gpu -x -q 2>&1 | tail -4
for S in 7 100 256; do PIXO_B200_SEGMENTS=$S timeout 600 python -m pytest tests/test_jpeg_gpu.py -m gpu -x -q -k "segment or dense or c2 or batch" 2>&1 | tail -1; done
timeout 300 python tools/prof_c4.py 4 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c4.csv python tools/prof_c4.py 2 > gpurun_out/prof_c4.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/launches_c4.csv') if not l.startswith('==')))
h=rows[0]; ci={n:i for i,n in enumerate(h)}
names=[(r[ci['Kernel Name']][:50], r[ci['Metric Value']]) for r in rows[1:] if len(r)>=len(h)]
for n in names[-6:]: print(n)
PY
timeout 900 python bench.py --configs C4 --steps 5 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_c4.json'):
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps(d['configs'])[:900])
PY
