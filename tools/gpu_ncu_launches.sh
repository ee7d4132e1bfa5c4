#!/bin/bash
# launch list of one timed bench step (ncu, durations only); run only after the plain command exits 0
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --iters-per-step 5"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_launches.log 2>&1
tail -n 2 gpurun_out/ncu_launches.log
