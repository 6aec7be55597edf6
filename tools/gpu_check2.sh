#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; t0=$SECONDS; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? after $((SECONDS-t0))s" | tee -a gpurun_out/$name.log; tail -n 6 gpurun_out/$name.log | cut -c1-600; }
run t_c3tc python -m pytest tests/test_kernels_gpu.py -q -x -k "c3_tc or masked"
run t_glue python -m pytest tests/test_kernels_gpu.py -q -k "not conv_down and not conv_up and not conv_wgrad and not c3_tc"
run t_tc python -m pytest tests/test_kernels_gpu.py -q -k "tc and (conv_down or conv_up or conv_wgrad)"
run t_model python -m pytest tests/test_model_gpu.py -q -x
run bench_graphs python bench.py --steps 30 --warmup 6
export DISCOGAN_B200_GRAPHS=0
run bench_eager python bench.py --steps 30 --warmup 6 --no-cpu-baseline --also-512 0
grep '^{' gpurun_out/bench_graphs.log > gpurun_out/bench_graphs.json
