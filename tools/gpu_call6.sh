#!/bin/bash
# round-2 call 6: coal-mining test, ncu launch list + ncu --set full of one launch of each CG-iteration kernel (C3)
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "coal" > gpurun_out/r2_pytest_coal.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_coal.log)
tail -n 4 gpurun_out/r2_pytest_coal.log | cut -c1-300
export MGVIB200_GRAPH=0
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --no-configs --iters-per-step 5"
$CMD > gpurun_out/r2_ncu_plain.json 2> gpurun_out/r2_ncu_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c3.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
tail -n 2 gpurun_out/r2_ncu_launches.log
$CMD > gpurun_out/r2_ncu_plain2.json 2> gpurun_out/r2_ncu_plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:"k_field_pass|k_vec" -s 60 -c 4 -o /tmp/prof_r2 $CMD > gpurun_out/r2_ncu_full.log 2>&1
ncu -i /tmp/prof_r2.ncu-rep --page raw --csv > gpurun_out/r2_prof_raw.csv 2>/dev/null
ncu -i /tmp/prof_r2.ncu-rep --page source --csv > gpurun_out/r2_prof_source.csv 2>/dev/null
tail -n 2 gpurun_out/r2_ncu_full.log; ls -la gpurun_out/r2_prof_raw.csv gpurun_out/r2_prof_source.csv
