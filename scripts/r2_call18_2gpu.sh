#!/bin/bash
# Round-2 GPU call 18 (2 GPUs): split expansion (SH columns beside the all-reduce) against the serial expansion, both exchange kinds
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29540
for v in "nccl_overlap:--exchange nccl" "nccl_serial:--exchange nccl --serial-expansion" "multimem_overlap:--exchange multimem" "multimem_serial:--exchange multimem --serial-expansion"; do
  name=${v%%:*}; flags=${v#*:}; port=$((port+1))
  echo "== bench N=2, $name"
  timeout 300 $RUN --master-port $port bench.py --gpus 2 --steps 20 --warmup 5 --repeats 4 $flags > gpurun_out/r2_bench_n2_$name.json 2> gpurun_out/r2_bench_n2_$name.err
  echo "rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n2_$name.err | tail -4
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r2_bench_n2_{name}.json") if l.startswith("{")][-1])
    ec = d.get("exchange_check", {})
    print(name, "value", d["value"], "ms_per_step", d["ms_per_step"], "spread", d.get("spread"), "err", ec.get("max_abs_err_over_max_abs_grad_features"),
          "noise", ec.get("dense_rerun_noise_features"), "dense_ms", ec.get("dense_all_reduce_ms_per_step"), "alone_ms", ec.get("compact_collectives_alone_ms"))
except Exception as e:
    print(name, "no line:", e)
PY
done
} 2>&1 | tee gpurun_out/r2_call18.log
