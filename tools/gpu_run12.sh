#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
for L in 2 4; do
MGVIB200_STRIDED_L=$L timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline > gpurun_out/sweep12_L$L.json 2> gpurun_out/sweep12_L$L.err
done
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --iters-per-step 5"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err && \
ncu --set full --clock-control none -k regex:"k_field_pass|k_vec" -s 60 -c 4 -o /tmp/prof_r1_final $CMD > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/prof_r1_final.ncu-rep --page raw --csv > gpurun_out/prof_r1_final2_raw.csv 2>/dev/null
tail -n 3 gpurun_out/pytest_gpu.log
