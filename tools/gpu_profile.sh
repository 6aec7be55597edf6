#!/bin/bash
# ncu evidence for profiles/: launch list of the default bench command, then --set full captures of the dominant kernels
# at both sizes.  Every ncu run follows a plain run of the same command that exited 0 (B200_PROFILING.md).
mkdir -p gpurun_out
B64="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-roofline --also-512 0 --no-inference --no-pipeline"
B512="python bench.py --image-size 512 --steps 3 --warmup 0 --no-cpu-baseline --no-roofline --no-inference --no-pipeline"
$B64 > gpurun_out/plain64.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_64.csv $B64 > gpurun_out/ncu_l64.log 2>&1
echo "launch list 64: $?"
$B512 > gpurun_out/plain512.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_512.csv $B512 > gpurun_out/ncu_l512.log 2>&1
echo "launch list 512: $?"
export DISCOGAN_B200_GRAPHS=0
ncu --set full --clock-control none --import-source on -k regex:"wgrad_gemm_pair_kernel|conv_gemm_kernel|conv_gemm_swap_kernel|bn_stream_kernel" -s 400 -c 60 -o gpurun_out/r02_full_512 $B512 > gpurun_out/ncu_f512.log 2>&1
echo "full 512: $?"
ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_kernel|conv_gemm_swap_kernel|wgrad_gemm" -s 600 -c 40 -o gpurun_out/r02_full_64 $B64 > gpurun_out/ncu_f64.log 2>&1
echo "full 64: $?"
ls -la gpurun_out/*.ncu-rep
