#!/bin/bash
mkdir -p gpurun_out
for L in 2 4; do
MGVIB200_STRIDED_L=$L timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline > gpurun_out/sweep13_L$L.json 2> gpurun_out/sweep13_L$L.err
done
(timeout 600 python -m pytest tests -m gpu -q -x -k "not long_1d and not c2_1d" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
tail -n 3 gpurun_out/pytest_gpu.log
