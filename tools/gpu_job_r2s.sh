#!/bin/bash
mkdir -p gpurun_out
for S in 64 128 256; do echo "== 16K frame, segments: $S"; PIXO_B200_SEGMENTS=$S timeout 300 python tools/prof_c4.py 4 2>&1 | tail -2; done | tee gpurun_out/seg_sweep_r2s.txt
PIXO_B200_SEGMENTS=256 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c4_s256.csv python tools/prof_c4.py 2 > gpurun_out/prof_c4.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/launches_c4_s256.csv') if not l.startswith('==')))
h=rows[0]; ci={n:i for i,n in enumerate(h)}
names=[(r[ci['Kernel Name']][:50], r[ci['Metric Value']]) for r in rows[1:] if len(r)>=len(h)]
for n in names[-6:]: print(n)
PY
