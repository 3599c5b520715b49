#!/bin/bash
# call 16: sort v2 (live-bit compaction, three buffers, early exit, shared-space addressing) + forward-blend bookkeeping
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== default"; timeout 300 python scripts/bench_stages.py C3
for v in lb8 lb2 mb4; do
  echo "== variant $v"; GSB200_LIB_PATH=$PWD/taichi_3d_gaussian_splatting_b200/libgsb_$v.so timeout 300 python scripts/bench_stages.py C3
done
echo "== C3s"; timeout 300 python scripts/bench_stages.py C3s
echo "== sort vs torch"; timeout 300 python scripts/bench_sort_vs_torch.py 2>&1 | tail -12
echo "== bench N=1"; timeout 600 python bench.py > gpurun_out/r2_bench_n1_call16.json 2> gpurun_out/r2_bench_n1_call16.err; echo rc=$?; head -c 600 gpurun_out/r2_bench_n1_call16.json
} 2>&1 | tee gpurun_out/r2_call16.log
