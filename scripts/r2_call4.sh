#!/bin/bash
# Round-2 GPU call 4: full GPU suite (exchange, fused step, controller epilogue), C5 harness in all modes
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
echo "== full GPU suite"; timeout 1500 python -m pytest -q -m gpu tests -x 2>&1 | tail -25
echo "== C5 autograd"; timeout 300 python scripts/train_c5.py 1000 2>/dev/null | tail -1
echo "== C5 fused kernels"; timeout 300 python scripts/train_c5.py 1000 --fused 2>/dev/null | tail -1
echo "== C5 fused step (Truck YAML: warm-up 1000)"; timeout 300 python scripts/train_c5.py 1000 --fused-step 2>/dev/null | tail -1
echo "== C5 fused step, densification from iteration 300, shuffled views, oracle parity over 200 iterations"
timeout 900 python scripts/train_c5.py 1000 --fused-step --warm-up 300 --shuffle --oracle-parity 200 2>gpurun_out/c5_parity.err | tail -1; tail -3 gpurun_out/c5_parity.err
echo "== C5 autograd, densification from iteration 300, shuffled views"; timeout 300 python scripts/train_c5.py 1000 --warm-up 300 --shuffle 2>/dev/null | tail -1
} 2>&1 | tee gpurun_out/r2_call4.log
