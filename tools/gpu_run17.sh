#!/bin/bash
mkdir -p gpurun_out
for T in 1 0; do
MGVIB200_TILED=$T timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline > gpurun_out/sweep17_T$T.json 2> gpurun_out/sweep17_T$T.err
done
MGVIB200_STRIDED_L=4 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline > gpurun_out/sweep17_T1_L4.json 2> gpurun_out/sweep17_T1_L4.err
(timeout 900 python -m pytest tests -m gpu -q -x -k "field_ops_vs_oracle or full_size_c3 or step_properties or lockstep" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
tail -n 3 gpurun_out/pytest_gpu.log
