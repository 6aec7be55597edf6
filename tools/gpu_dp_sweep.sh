#!/bin/bash
# 8-GPU same-box sweep of the gradient-exchange knobs at 64x64 (each line: env settings, ms/step, pairs/s)
N=${1:-8}
mkdir -p gpurun_out
i=0
while read -r envs; do
  i=$((i+1))
  env $envs timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+i)) \
    bench.py --gpus $N --steps 120 --warmup 12 --no-cpu-baseline --also-512 ${ALSO:-0} > gpurun_out/sweep_$i.log 2> gpurun_out/sweep_$i.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/sweep_$i.log") if l.startswith("{")][0]
    a = d.get("also") or {}
    print("$envs -> %.4f ms/step %.0f pairs/s e2e %.0f | 512: %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], a.get("ms_per_step")))
except Exception as e:
    print("$envs failed", e, open("gpurun_out/sweep_$i.err").read()[-400:])
PY
done <<LIST
X=0
NCCL_MIN_CTAS=16
NCCL_MIN_CTAS=32
NCCL_ALGO=Ring
NCCL_ALGO=Tree
LIST
