#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -q > gpurun_out/t_model.log 2>&1; echo "model tests exit $?"; tail -n 5 gpurun_out/t_model.log
timeout 900 python bench.py --steps 30 --warmup 6 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"; cat gpurun_out/bench_full.json; tail -n 5 gpurun_out/bench_full.err
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --also-512 0 --no-roofline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_64.csv $CMD > gpurun_out/ncu64.log 2>&1; echo "ncu64 exit $?"
CMD5="python bench.py --image-size 512 --steps 3 --warmup 3 --no-cpu-baseline --no-roofline"
timeout 600 $CMD5 > gpurun_out/plain512.log 2>&1 && timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1500 --csv --log-file gpurun_out/launches_512.csv $CMD5 > gpurun_out/ncu512.log 2>&1; echo "ncu512 exit $?"
cat gpurun_out/plain512.log | tail -n 3
