#!/bin/bash
# compute-sanitizer over the kernel tour (tools/sanitize_run.py): memcheck, then racecheck
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/san_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/san_racecheck.log
timeout 300 python -m pytest tests/test_png_gpu.py -m gpu -x -q 2>&1 | tail -2
