#!/bin/bash
# GPU session 2: correctness, then tuning sweep of resident blocks / strided tile width at C3.
mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log)
(timeout 1200 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
for lib in default mb2 mb4; do
  for L in 1 2 4; do
    if [ "$lib" = "default" ]; then unset MGVIB200_LIB; else export MGVIB200_LIB=$PWD/build/variants/libmgvib200_$lib.so; fi
    MGVIB200_STRIDED_L=$L timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline \
        > gpurun_out/sweep_${lib}_L$L.json 2> gpurun_out/sweep_${lib}_L$L.err
    echo "rc=$?" >> gpurun_out/sweep_${lib}_L$L.err
  done
done
unset MGVIB200_LIB
(timeout 900 python bench.py --steps 3 --warmup 3 --e2e-steps 1 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "rc=$?" >> gpurun_out/bench_c3.err)
tail -n 3 gpurun_out/smoke.log gpurun_out/pytest_gpu.log
