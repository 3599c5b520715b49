#!/bin/bash
# Round-2 GPU call 8 (2 GPUs): the hand-written NVLS exchange kernel against the NCCL collectives
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== bench N=2, multimem exchange"
timeout 600 $RUN --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 --repeats 3 --exchange multimem > gpurun_out/r2_bench_n2_multimem.json 2> gpurun_out/r2_bench_n2_multimem.err
echo "rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n2_multimem.err | tail -25; grep '^{' gpurun_out/r2_bench_n2_multimem.json | cut -c1-300
echo "== bench N=2, NCCL exchange"
timeout 600 $RUN --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 --repeats 3 --exchange nccl > gpurun_out/r2_bench_n2_nccl.json 2> gpurun_out/r2_bench_n2_nccl.err
echo "rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n2_nccl.err | tail -5; grep '^{' gpurun_out/r2_bench_n2_nccl.json | cut -c1-300
} 2>&1 | tee gpurun_out/r2_call8.log
