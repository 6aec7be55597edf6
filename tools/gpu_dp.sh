#!/bin/bash
# Data-parallel validation on N GPUs of one box (gpurun --gpus N): the DP-vs-sharded-oracle test, then the bench at 1 and N.
N=${1:-2}
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; t0=$SECONDS; timeout -k 5 ${T:-600} "$@" > gpurun_out/$name.log 2> gpurun_out/$name.err; echo "exit $? after $((SECONDS-t0))s"; tail -n 4 gpurun_out/$name.log | cut -c1-600; tail -n 3 gpurun_out/$name.err | cut -c1-300; }
[ -z "$SKIP_TESTS" ] && run t_dp python -m pytest tests/test_dp_gpu.py -m gpu -q
run bench_n1 python bench.py --steps 150 --warmup 15 --no-cpu-baseline --no-roofline --no-inference
run bench_n$N python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 150 --warmup 15 --no-cpu-baseline
