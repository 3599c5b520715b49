#!/bin/bash
# Round-2 GPU call 6 (8 GPUs): bench N=8 with the compact exchange (sequential / concurrent collectives), NCCL algorithm info
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
nvidia-smi -L | head -8
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "== bench N=8, compact exchange, one communicator"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 900 $RUN --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --repeats 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
grep -E "NVLS|Connected all|Channel 00/|nChannels|Algo" gpurun_out/r2_bench_n8.err | sort | uniq -c | sort -rn | head -12; grep -v NCCL gpurun_out/r2_bench_n8.err | tail -5; grep '^{' gpurun_out/r2_bench_n8.json | cut -c1-300
echo "== bench N=8, compact exchange, all-gather on a second communicator"
timeout 900 $RUN --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 --repeats 3 --exchange-streams 2 > gpurun_out/r2_bench_n8_2streams.json 2> gpurun_out/r2_bench_n8_2streams.err; tail -3 gpurun_out/r2_bench_n8_2streams.err; grep '^{' gpurun_out/r2_bench_n8_2streams.json | cut -c1-300
} 2>&1 | tee gpurun_out/r2_call6.log
