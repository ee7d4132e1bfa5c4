#!/bin/bash
mkdir -p gpurun_out
for L in 2 4; do
MGVIB200_STRIDED_L=$L timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline > gpurun_out/sweep18_L$L.json 2> gpurun_out/sweep18_L$L.err
done
MGVIB200_TILED=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline > gpurun_out/sweep18_T0.json 2> gpurun_out/sweep18_T0.err
(timeout 600 python -m pytest tests -m gpu -q -x -k "full_size_c3 or step_properties or (field_ops_vs_oracle and 2048)" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
tail -n 3 gpurun_out/pytest_gpu.log
