#!/bin/bash
# Final evidence of the round (1 GPU): bench line, ncu launch list of 2 steps, ncu --set full of every kernel of a step
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench.err; tail -2 gpurun_out/r02_bench.err; cut -c1-260 gpurun_out/r02_bench_line.json
echo "== ncu launch list (2 steps)"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python scripts/profile_step.py --steps 2 > gpurun_out/r02_launches.log 2>&1; tail -1 gpurun_out/r02_launches.log
python scripts/launch_shares.py gpurun_out/r02_launches.csv "ncu launch list of 2 fwd+bwd steps at C3" | tee gpurun_out/r02_launch_shares.txt
echo "== ncu --set full, every kernel of one step"
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'gsb' -o gpurun_out/r02_step_full python scripts/profile_step.py --steps 1 > gpurun_out/r02_full.log 2>&1; tail -2 gpurun_out/r02_full.log
} 2>&1 | tee gpurun_out/r2_final_profile.log
