#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; tmo=$2; shift; shift; echo "=== $name"; t0=$SECONDS; timeout -k 5 $tmo "$@" > gpurun_out/$name.log 2>&1; echo "exit $? after $((SECONDS-t0))s" | tee -a gpurun_out/$name.log; tail -n 3 gpurun_out/$name.log | cut -c1-300; }
run bench_dp4 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29565 bench.py --gpus 4 --steps 30 --warmup 6
run t_dp 200 python -m pytest tests/test_dp_gpu.py -q
run bench_dp4_512 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29566 bench.py --gpus 4 --steps 12 --warmup 6 --image-size 512
run bench_ref4 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29567 bench.py --impl reference --gpus 4 --steps 3 --warmup 1
run bench512 200 python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline
grep '^{' gpurun_out/bench512.log > gpurun_out/bench512.json
