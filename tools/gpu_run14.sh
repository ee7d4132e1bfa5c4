#!/bin/bash
mkdir -p gpurun_out
for S in 1 0; do
MGVIB200_STAGE_VIN=$S timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline > gpurun_out/sweep14_S$S.json 2> gpurun_out/sweep14_S$S.err
done
(timeout 900 python -m pytest tests -m gpu -q -x -k "not long_1d and not c2_1d and not full_size" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
tail -n 3 gpurun_out/pytest_gpu.log
