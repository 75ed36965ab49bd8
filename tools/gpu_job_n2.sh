#!/bin/bash
# two-GPU validation: the NCCL test of the tiled frame and a short N=2 bench (configs included)
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -8
N=${1:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1]   # NCCL may print a version line first
print('N',d['n_gpus'],'value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'single',d['e2e']['single_call_pageable']['value'])
for k,v in d['configs'].items(): print(k, json.dumps(v)[:1200])
PY
tail -5 gpurun_out/bench_n$N.err
