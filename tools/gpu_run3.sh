#!/bin/bash
# GPU session 3: failing tests rerun, then ncu launch list + full capture of the pass kernels at C3.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -k "default_tolerance or lockstep or polyfit" > gpurun_out/pytest_gpu_rerun.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_rerun.log)
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --iters-per-step 5"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 66 -c 22 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_field_pass -s 48 -c 3 -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/pytest_gpu_rerun.log gpurun_out/ncu_list.log gpurun_out/ncu_full.log
