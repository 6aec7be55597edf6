#!/bin/bash
# Same-box A/B of an environment switch: bash tools/gpu_ab.sh VAR "0 1" [bench args...]
VAR=$1; VALUES=$2; shift 2
mkdir -p gpurun_out
for v in $VALUES; do
  for S in 64 512; do
    env $VAR=$v timeout -k 5 300 python bench.py --image-size $S --steps $([ $S = 64 ] && echo 300 || echo 18) --warmup 12 \
      --no-cpu-baseline --no-roofline --no-inference --no-pipeline --also-512 0 "$@" > gpurun_out/ab_${VAR}_${v}_$S.json 2> gpurun_out/ab_${VAR}_${v}_$S.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_${VAR}_${v}_$S.json"))
    print("$VAR=$v S=$S: %.3f ms/step  %.0f pairs/s  launches/step %.0f" % (d["ms_per_step"], d["value"], d["gpu_launches"] / d["steps"]))
except Exception as e:
    print("$VAR=$v S=$S failed:", e); print(open("gpurun_out/ab_${VAR}_${v}_$S.err").read()[-600:])
PY
  done
done
