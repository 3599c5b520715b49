#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "== new tests"; timeout 600 python -m pytest -q -m gpu tests/test_gpu_parity.py tests/test_gpu_work_counters.py 2>&1 | tail -4
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_call11.json 2> gpurun_out/r2_bench_call11.err; tail -2 gpurun_out/r2_bench_call11.err; cut -c1-200 gpurun_out/r2_bench_call11.json
} 2>&1 | tee gpurun_out/r2_call11.log
