#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; t0=$SECONDS; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; echo "exit $? after $((SECONDS-t0))s" | tee -a gpurun_out/$name.log; tail -n 3 gpurun_out/$name.log | cut -c1-300; }
DISCOGAN_B200_FUSE_STATS=0 run bench512_nofuse python bench.py --image-size 512 --steps 12 --warmup 6 --no-cpu-baseline
grep '^{' gpurun_out/bench512_nofuse.log > gpurun_out/bench512_nofuse.json
run plain512 python tools/profile_cycle.py 512
run ncu512 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_512_cycle.csv python tools/profile_cycle.py 512
run plain64 python tools/profile_cycle.py 64
run ncu64 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_64_cycle.csv python tools/profile_cycle.py 64
