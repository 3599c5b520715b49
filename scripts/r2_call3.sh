#!/bin/bash
# Round-2 GPU call 3 (2 GPUs): compact exchange test, bench N=1 and N=2
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
nvidia-smi -L
echo "== exchange + counters tests"; timeout 600 python -m pytest -q -m gpu tests/test_gpu_exchange.py tests/test_gpu_work_counters.py 2>&1 | tail -5
echo "== bench N=1"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -3 gpurun_out/r2_bench_n1.err; cut -c1-300 gpurun_out/r2_bench_n1.json
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -5 gpurun_out/r2_bench_n2.err; cut -c1-300 gpurun_out/r2_bench_n2.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | tee gpurun_out/r2_bench_ref.json | cut -c1-300
} 2>&1 | tee gpurun_out/r2_call3.log
