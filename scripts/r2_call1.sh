#!/bin/bash
# Round-2 GPU call 1: pending GPU tests, new full-size parity tests (counts recorded), experimental kernels, stage times
# of both backward kernels, sort vs torch.sort, bench lines, C5, one ncu capture of the transposed kernel.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv
echo "== full-size parity (C2, C3, C3s; both paths)"; timeout 1200 python -m pytest -q -m gpu tests/test_gpu_full_size_parity.py 2>&1 | tail -30
echo "== golden + behaviour + large splats"; timeout 900 python -m pytest -q -m gpu tests/test_reference_behaviour.py tests/test_reference_path_golden.py tests/test_gpu_zz_large_splats.py 2>&1 | tail -15
echo "== experimental backward (runxfail)"; timeout 600 python -m pytest -q -m gpu --runxfail tests/test_zzz_experimental_gpu.py 2>&1 | tail -25
for wl in C3 C2; do
  echo "== stages $wl default"; timeout 300 python scripts/bench_stages.py $wl
  echo "== stages $wl default kernel without hook statistics"; GSB200_SKIP_HOOK_STATS=1 timeout 300 python scripts/bench_stages.py $wl
  echo "== stages $wl transposed"; GSB200_BACKWARD_IMPL=transposed timeout 300 python scripts/bench_stages.py $wl
done
echo "== sort vs torch"; timeout 300 python scripts/bench_sort_vs_torch.py
echo "== bench default"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline | tee gpurun_out/r2_bench_default.json | cut -c1-400
echo "== bench transposed"; GSB200_BACKWARD_IMPL=transposed timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline | tee gpurun_out/r2_bench_transposed.json | cut -c1-400
echo "== C5 trainer, torch step vs fused step"; timeout 300 python scripts/train_c5.py 1000; timeout 300 python scripts/train_c5.py 1000 --fused
echo "== ncu transposed kernel"
GSB200_BACKWARD_IMPL=transposed timeout 600 ncu --set full --clock-control none --import-source on -k regex:blend_backward_transposed -s 2 -c 1 \
  -o gpurun_out/r2_bwd_transposed python scripts/bench_stages.py C3 > gpurun_out/r2_ncu.log 2>&1; tail -3 gpurun_out/r2_ncu.log
} 2>&1 | tee gpurun_out/r2_call1.log
