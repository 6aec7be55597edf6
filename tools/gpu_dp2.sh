#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; tmo=$2; shift; shift; echo "=== $name"; t0=$SECONDS; timeout -k 5 $tmo "$@" > gpurun_out/$name.log 2>&1; echo "exit $? after $((SECONDS-t0))s" | tee -a gpurun_out/$name.log; tail -n 5 gpurun_out/$name.log | cut -c1-400; }
run bench_dp2 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 90 --warmup 9
run t_dp 240 python -m pytest tests/test_dp_gpu.py -q -x
run bench_dp2_512 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 12 --warmup 6 --image-size 512
