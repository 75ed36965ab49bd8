#!/bin/bash
# round-2 final single-GPU run: parity suite, smoke, the bench line, the ncu launch list of the same bench
# command, one --set full capture per hot kernel, and a launch list of the 16K-frame path
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'single',d['e2e']['single_call_pageable']['ms_per_call'],'launches',d['gpu_launches'],'roof',d['roofline']['frac'], d['roofline'].get('entropy_kernel_ms_per_step'), d['clocks'])
print('cpu', d['cpu_baseline'])
for k,v in d['configs'].items(): print(k, json.dumps(v)[:700])
PY
tail -3 gpurun_out/bench_final.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; head -c 600 gpurun_out/bench_ref.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --configs none > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c4.csv python tools/prof_c4.py 2 > gpurun_out/prof_c4.log 2>&1
tail -3 gpurun_out/prof_c4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_png_band -s 1 -c 1 -f -o gpurun_out/png_final2 python tools/prof_run.py png 16 3 > gpurun_out/ncu_png.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_huff -s 2 -c 1 -f -o gpurun_out/huff_final2 python tools/prof_run.py encode 32 4 > gpurun_out/ncu_huff.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_jpeg_420 -s 2 -c 1 -f -o gpurun_out/k1_final2 python tools/prof_run.py encode 32 4 > gpurun_out/ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_jpeg_444 -s 1 -c 1 -f -o gpurun_out/k444_final2 python tools/prof_444.py > gpurun_out/ncu_k444.log 2>&1
for f in png huff k1 k444; do tail -n 1 gpurun_out/ncu_$f.log; done
ls -la gpurun_out/*final2* gpurun_out/launches_*.csv
