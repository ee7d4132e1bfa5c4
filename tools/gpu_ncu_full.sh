#!/bin/bash
# ncu --set full of one launch of each CG-iteration kernel; the report stays on the box, the raw page comes back
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --iters-per-step 5"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:"k_field_pass|k_vec" -s 60 -c 4 -o /tmp/prof_final $CMD > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_final_raw.csv 2>/dev/null
tail -n 2 gpurun_out/ncu_full.log
