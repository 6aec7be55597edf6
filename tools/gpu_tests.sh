#!/bin/bash
# Run the GPU test-suite (or a subset: arguments are passed to pytest) on the box; the log and the parity report land
# in gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
if [ $# -eq 0 ]; then set -- tests; fi
timeout -k 5 ${DG_TEST_TIMEOUT:-1500} python -m pytest "$@" -m gpu -q --durations=8 --tb=short 2>&1 | cut -c1-400 | tail -n 400 > gpurun_out/t_gpu.log
tail -n 25 gpurun_out/t_gpu.log
