#!/bin/bash
# Round-2 GPU call 7: TMA-staged preprocess + incremental phase 2 on the GPU: parity tests, stage times, variants
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "== parity tests"; timeout 1500 python -m pytest -q -m gpu tests/test_gpu_parity.py tests/test_gpu_full_size_parity.py tests/test_reference_path_golden.py tests/test_gpu_zz_large_splats.py tests/test_zzz_experimental_gpu.py tests/test_gpu_fused_step.py tests/test_reference_behaviour.py 2>&1 | tail -8
for wl in C3 C2 C3s; do echo "== stages $wl"; timeout 300 python scripts/bench_stages.py $wl; done
for v in pre4 scan256 scan64; do
  echo "== variant $v"; GSB200_LIB_PATH=$PWD/taichi_3d_gaussian_splatting_b200/libgsb_$v.so timeout 300 python scripts/bench_stages.py C3
done
echo "== ncu preprocess + sort pass"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'preprocess_kernel|onesweep_pass' -s 5 -c 5 \
  -o gpurun_out/r2_pre_sort python scripts/bench_stages.py C3 > gpurun_out/r2_ncu7.log 2>&1; tail -2 gpurun_out/r2_ncu7.log
} 2>&1 | tee gpurun_out/r2_call7.log
