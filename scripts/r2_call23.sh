#!/bin/bash
# call 23 (1 GPU, the round's last GPU seconds): per-point kernel with the index-only loads hoisted (variant build) against the default
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
V=$PWD/taichi_3d_gaussian_splatting_b200/libgsb_hoist.so
{
echo "== default"; timeout 100 python scripts/bench_stages.py C3
echo "== variant hoist"; GSB200_LIB_PATH=$V timeout 100 python scripts/bench_stages.py C3
echo "== variant: parity tests"; GSB200_LIB_PATH=$V timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_reference_path_golden.py -m gpu -x -q 2>&1 | tail -3
echo "== variant: full-size parity"; GSB200_LIB_PATH=$V timeout 300 python -m pytest tests/test_gpu_full_size_parity.py -m gpu -x -q 2>&1 | tail -3
} 2>&1 | tee gpurun_out/r2_call23.log
