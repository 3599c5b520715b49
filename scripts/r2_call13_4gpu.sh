#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
echo "== bench N=4 (default exchange = auto)"
timeout 600 $RUN --master-port 29551 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
echo "rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n4.err | tail -6; grep '^{' gpurun_out/r2_bench_n4.json | cut -c1-200
echo "== reference arm under torchrun N=4"
timeout 600 $RUN --master-port 29552 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 | cut -c1-200
} 2>&1 | tee gpurun_out/r2_call13.log
