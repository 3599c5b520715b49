#!/bin/bash
# Round-2 GPU call 5: the whole GPU suite (no -x), timing of the reference-size stress test
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "== full GPU suite"; timeout 2400 python -m pytest -q -m gpu tests --durations=8 2>&1 | tail -40
} 2>&1 | tee gpurun_out/r2_call5.log
