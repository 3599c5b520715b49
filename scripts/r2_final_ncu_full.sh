#!/bin/bash
# ncu --set full (with source) of every kernel of one fwd+bwd step at C3 (the -k filter matches the BASE kernel name: no namespace)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
timeout 800 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'kernel' -o gpurun_out/r02_step_full python scripts/profile_step.py --steps 1 > gpurun_out/r02_full.log 2>&1; tail -3 gpurun_out/r02_full.log; ls -la gpurun_out/r02_step_full.ncu-rep
} 2>&1 | tee gpurun_out/r2_final_ncu_full.log
