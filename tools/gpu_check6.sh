#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; t0=$SECONDS; timeout 1500 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? after $((SECONDS-t0))s" | tee -a gpurun_out/$name.log; tail -n 3 gpurun_out/$name.log | cut -c1-300; }
run t_kernels python -m pytest tests/test_kernels_gpu.py -q -k "not simt"
run t_model python -m pytest tests/test_model_gpu.py -q
run bench64 python bench.py --steps 30 --warmup 6 --also-512 0
run bench512 python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline
grep '^{' gpurun_out/bench64.log > gpurun_out/bench64.json
grep '^{' gpurun_out/bench512.log > gpurun_out/bench512.json
run plain512 python tools/profile_cycle.py 512
run ncu_gemm ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm_kernel -c 13 -o gpurun_out/prof512_conv_gemm -f python tools/profile_cycle.py 512
run ncu_c3 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:c3_down_tc_kernel -c 1 -o gpurun_out/prof512_c3_down -f python tools/profile_cycle.py 512
run ncu_wgrad ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:wgrad_gemm_kernel -c 6 -o gpurun_out/prof512_wgrad -f python tools/profile_cycle.py 512
ls -la gpurun_out/*.ncu-rep
