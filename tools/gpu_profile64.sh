#!/bin/bash
# ncu --set full of the dominant kernels of the eager 64x64 step (a plain run of the same command first)
mkdir -p gpurun_out
export DISCOGAN_B200_GRAPHS=0
B64="python bench.py --steps 3 --warmup 0 --no-cpu-baseline --no-roofline --also-512 0 --no-inference --no-pipeline"
$B64 > gpurun_out/plain64.log 2>&1 && \
timeout -k 5 420 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm|wgrad_gemm" -s 200 -c 40 -o gpurun_out/r02_full_64 $B64 > gpurun_out/ncu_f64.log 2>&1
echo "full 64: $?"; ls -la gpurun_out/r02_full_64.ncu-rep
