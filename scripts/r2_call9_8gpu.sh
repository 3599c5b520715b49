#!/bin/bash
# Round-2 GPU call 9 (8 GPUs): hand-written NVLS exchange kernel vs the NCCL collectives at 8 ranks
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
i=0
for v in "multimem:--exchange multimem" "multimem_b592:--exchange multimem --exchange-blocks 592" "nccl:--exchange nccl"; do
  name=${v%%:*}; flags=${v#*:}; i=$((i+1))
  echo "== bench N=8, $name"
  timeout 600 $RUN --master-port $((29540+i)) bench.py --gpus 8 --steps 20 --warmup 5 --repeats 5 $flags > gpurun_out/r2_bench_n8_$name.json 2> gpurun_out/r2_bench_n8_$name.err
  echo "rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n8_$name.err | tail -6; grep '^{' gpurun_out/r2_bench_n8_$name.json | cut -c1-200
done
} 2>&1 | tee gpurun_out/r2_call9.log
