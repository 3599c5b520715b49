#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "== bench N=8 (default: auto -> multimem with the early push)"
timeout 600 $RUN --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_final.json 2> gpurun_out/r2_bench_n8_final.err
echo "rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n8_final.err | tail -6; grep '^{' gpurun_out/r2_bench_n8_final.json | cut -c1-200
} 2>&1 | tee gpurun_out/r2_call15.log
