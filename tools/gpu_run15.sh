#!/bin/bash
mkdir -p gpurun_out
for L in 0 2 4; do
  if [ $L = 0 ]; then unset MGVIB200_STRIDED_L; else export MGVIB200_STRIDED_L=$L; fi
  timeout 400 python bench.py --workload c5 --steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline > gpurun_out/sweep15_c5_L$L.json 2> gpurun_out/sweep15_c5_L$L.err
done
unset MGVIB200_STRIDED_L
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline > gpurun_out/sweep15_c3.json 2> gpurun_out/sweep15_c3.err
(timeout 600 python -m pytest tests -m gpu -q -x -k "field_ops_vs_oracle or full_size_c5 or metric_1d" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
tail -n 3 gpurun_out/pytest_gpu.log
