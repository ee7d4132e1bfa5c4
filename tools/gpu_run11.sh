#!/bin/bash
mkdir -p gpurun_out
for PE in 0 1; do
  MGVIB200_PREFETCH_EPI=$PE timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline > gpurun_out/sweep9_pe$PE.json 2> gpurun_out/sweep9_pe$PE.err
done
( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2> gpurun_out/bench_default.time
(timeout 300 python bench.py --workload c5 --steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --iters-per-step 10 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "rc=$?" >> gpurun_out/bench_c5.err)
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --iters-per-step 5"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 66 -c 22 --csv --log-file gpurun_out/launches_r1_final.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none -k regex:"k_field_pass|k_vec" -s 60 -c 4 -o /tmp/prof_r1_final $CMD > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/prof_r1_final.ncu-rep --page raw --csv > gpurun_out/prof_r1_final_raw.csv 2>/dev/null
ls -la /tmp/prof_r1_final.ncu-rep >> gpurun_out/ncu_full.log
du -sh gpurun_out; cat gpurun_out/bench_default.time
