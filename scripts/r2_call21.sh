#!/bin/bash
# call 21 (1 GPU): SH colour in a TMA-fed kernel of its own, forked beside the radix sort -- measured SLOWER and reverted
# (profiles/r02_call21_sh_color_split.log; this script ran against that build)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== stages (serial launches)"; timeout 300 python scripts/bench_stages.py C3
echo "== step, colour kernel beside the sort vs serial"; timeout 300 python scripts/bench_step.py C3
echo "== C2"; timeout 300 python scripts/bench_step.py C2
} 2>&1 | tee gpurun_out/r2_call21.log
