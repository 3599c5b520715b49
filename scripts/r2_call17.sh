#!/bin/bash
# call 17: vectorised histogram sweep, split expansion (emulated ranks on one GPU), launch list of the new sort
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== default"; timeout 300 python scripts/bench_stages.py C3
echo "== ncu launch list (2 steps)"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_call17_launches.csv python scripts/profile_step.py --steps 2 > gpurun_out/r2_call17_launches.log 2>&1; tail -1 gpurun_out/r2_call17_launches.log
python scripts/launch_shares.py gpurun_out/r2_call17_launches.csv "ncu launch list of 2 fwd+bwd steps at C3 (call 17)" | tee gpurun_out/r2_call17_launch_shares.txt
} 2>&1 | tee gpurun_out/r2_call17.log
