#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "== default"; timeout 300 python scripts/bench_stages.py C3
for v in p1u4 p1u1 fwu8 fwu2 fwu8b4; do
  echo "== variant $v"; GSB200_LIB_PATH=$PWD/taichi_3d_gaussian_splatting_b200/libgsb_$v.so timeout 300 python scripts/bench_stages.py C3
done
} 2>&1 | tee gpurun_out/r2_call12.log
