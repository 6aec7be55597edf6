#!/bin/bash
# First-pass GPU check: each group in its own process so a trapped kernel cannot poison the others.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" | tee -a gpurun_out/$name.log; tail -n 15 gpurun_out/$name.log; }
run t_glue python -m pytest tests/test_kernels_gpu.py -q -k "not conv_down and not conv_up and not conv_wgrad"
run t_simt python -m pytest tests/test_kernels_gpu.py -q -k "simt"
run diag_down python tools/diag_tc.py down
run diag_up python tools/diag_tc.py up
run diag_wgrad python tools/diag_tc.py wgrad
run t_tc_down python -m pytest tests/test_kernels_gpu.py -q -k "conv_down and tc"
run t_tc_up python -m pytest tests/test_kernels_gpu.py -q -k "conv_up and tc"
run t_tc_wgrad python -m pytest tests/test_kernels_gpu.py -q -k "conv_wgrad and tc"
run t_model python -m pytest tests/test_model_gpu.py -q
