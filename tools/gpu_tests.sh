#!/bin/bash
# Run the GPU test-suite (or a subset: args are passed to pytest) on the box; logs + the parity report land in gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
args=${@:-tests}
timeout -k 5 ${DG_TEST_TIMEOUT:-1500} python -m pytest $args -m gpu -q -x --durations=8 2>&1 | tail -n 80 > gpurun_out/t_gpu.log
tail -n 25 gpurun_out/t_gpu.log
