#!/bin/bash
# word-wise splice kernels: parity (segments, bands, C4 full size) and the 16K-frame launch list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_jpeg_gpu.py tests/test_multi_gpu.py tests/test_configs_full_gpu.py tests/test_golden_reference.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/prof_c4.py 4 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c4.csv python tools/prof_c4.py 2 > gpurun_out/prof_c4.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/launches_c4.csv') if not l.startswith('==')))
h=rows[0]; ci={n:i for i,n in enumerate(h)}
for r in rows[1:]:
    if len(r)>=len(h): last=r
names=[(r[ci['Kernel Name']][:50], r[ci['Metric Value']]) for r in rows[1:] if len(r)>=len(h)]
for n in names[-6:]: print(n)
PY
