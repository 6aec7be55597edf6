#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --also-512 0 --no-roofline"
timeout 200 $CMD > gpurun_out/plain64.log 2>&1 || exit 1
timeout -k 5 300 ncu --set full --clock-control none -k regex:conv_gemm -s 40 -c 20 -o /tmp/ncu/prof64 -f $CMD > gpurun_out/ncu64_full.log 2>&1; echo "ncu exit $?"
timeout 120 ncu -i /tmp/ncu/prof64.ncu-rep --page raw --csv > gpurun_out/ncu64_conv_gemm_raw.csv 2>/dev/null; ls -la gpurun_out/ncu64_conv_gemm_raw.csv
