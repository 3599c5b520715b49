#!/bin/bash
# Round-2 GPU call 2: full GPU suite on the rewritten forward blend / unified alpha, stage times + variants, ncu of the forward.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "== full GPU suite"; timeout 1500 python -m pytest -q -m gpu tests 2>&1 | tail -15
for wl in C3 C2 C3s; do
  echo "== stages $wl"; timeout 300 python scripts/bench_stages.py $wl
done
echo "== stages C3 butterfly"; GSB_BENCH_BACKWARD_IMPL=butterfly timeout 300 python scripts/bench_stages.py C3
for v in fwd4 fwd6 tb2; do
  echo "== variant $v"; GSB200_LIB_PATH=$PWD/taichi_3d_gaussian_splatting_b200/libgsb_$v.so timeout 300 python scripts/bench_stages.py C3
done
echo "== bench"; timeout 600 python bench.py --steps 30 --warmup 5 | tee gpurun_out/r2_bench_call2.json | cut -c1-600
echo "== ncu forward + backward kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'blend_forward_kernel|blend_backward_transposed' -s 4 -c 2 \
  -o gpurun_out/r2_blend python scripts/bench_stages.py C3 > gpurun_out/r2_ncu2.log 2>&1; tail -3 gpurun_out/r2_ncu2.log
} 2>&1 | tee gpurun_out/r2_call2.log
