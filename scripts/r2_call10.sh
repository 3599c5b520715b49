#!/bin/bash
# Round-2 GPU call 10: whole GPU suite on the current tree, bench N=1 (median-region headline), C5 harness with a warm-up run
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "== full GPU suite"; timeout 1500 python -m pytest -q -m gpu tests 2>&1 | tail -6
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_call10.json 2> gpurun_out/r2_bench_call10.err; tail -2 gpurun_out/r2_bench_call10.err; cut -c1-260 gpurun_out/r2_bench_call10.json
echo "== C5 fused step (Truck YAML)"; timeout 300 python scripts/train_c5.py 1000 --fused-step 2>/dev/null | tail -1
echo "== C5 fused step, warm-up 300, shuffled"; timeout 300 python scripts/train_c5.py 1000 --fused-step --warm-up 300 --shuffle 2>/dev/null | tail -1
echo "== C5 autograd (Truck YAML)"; timeout 300 python scripts/train_c5.py 1000 2>/dev/null | tail -1
} 2>&1 | tee gpurun_out/r2_call10.log
