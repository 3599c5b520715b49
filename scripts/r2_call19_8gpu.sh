#!/bin/bash
# Round-2 GPU call 19 (8 GPUs): NVLS exchange with the SH expansion beside the all-reduce (default) vs the serial expansion
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
port=29570
for v in "overlap:" "serial:--serial-expansion --repeats 4" "overlap_b148:--exchange-blocks 148 --repeats 4"; do
  name=${v%%:*}; flags=${v#*:}; port=$((port+1))
  echo "== bench N=8, $name"
  timeout 240 $RUN --master-port $port bench.py --gpus 8 --steps 20 --warmup 5 $flags > gpurun_out/r2_bench_n8_$name.json 2> gpurun_out/r2_bench_n8_$name.err
  echo "rc=$?"; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n8_$name.err | tail -4
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/r2_bench_n8_{name}.json") if l.startswith("{")][-1])
    ec = d.get("exchange_check", {})
    print(name, "value", d["value"], "ms_per_step", d["ms_per_step"], "spread", {k: v for k, v in d.get("spread", {}).items() if k != "what"},
          "err", ec.get("max_abs_err_over_max_abs_grad_features"), "noise", ec.get("dense_rerun_noise_features"),
          "dense_ms", ec.get("dense_all_reduce_ms_per_step"), "alone_ms", ec.get("compact_collectives_alone_ms"),
          "per_rank", ec.get("fwd_bwd_without_exchange_ms_per_rank"), "C4", (d.get("other_configs") or {}).get("C4"))
except Exception as e:
    print(name, "no line:", e)
PY
done
} 2>&1 | tee gpurun_out/r2_call19.log
