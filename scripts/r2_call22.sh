#!/bin/bash
# call 22 (1 GPU): radix ranking with the MATCH.ANY loop split off (variant build) against the default
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
V=$PWD/taichi_3d_gaussian_splatting_b200/libgsb_mf1.so
{
echo "== variant: sort + parity tests"; GSB200_LIB_PATH=$V timeout 600 python -m pytest tests/test_gpu_sort.py tests/test_gpu_parity.py tests/test_reference_path_golden.py -m gpu -x -q 2>&1 | tail -3
echo "== default"; timeout 300 python scripts/bench_stages.py C3
echo "== variant mf1"; GSB200_LIB_PATH=$V timeout 300 python scripts/bench_stages.py C3
echo "== default again"; timeout 300 python scripts/bench_stages.py C3
echo "== variant mf1 again"; GSB200_LIB_PATH=$V timeout 300 python scripts/bench_stages.py C3
echo "== C4 scene, default"; timeout 300 python scripts/bench_stages.py C4
} 2>&1 | tee gpurun_out/r2_call22.log
