#!/bin/bash
# call 20 (1 GPU): 64-bit (G, alpha T) exchange in the backward blend; occupancy / grid knobs; then the round's final evidence
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== default"; timeout 300 python scripts/bench_stages.py C3
for v in hb2 hb8 pre6 pre5 fwb5; do
  echo "== variant $v"; GSB200_LIB_PATH=$PWD/taichi_3d_gaussian_splatting_b200/libgsb_$v.so timeout 300 python scripts/bench_stages.py C3
done
} 2>&1 | tee gpurun_out/r2_call20.log
