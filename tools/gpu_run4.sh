#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log)
(timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
for lib in default mb2 mb4; do
  for L in 2 4; do
    if [ "$lib" = "default" ]; then unset MGVIB200_LIB; else export MGVIB200_LIB=$PWD/build/variants/libmgvib200_$lib.so; fi
    MGVIB200_STRIDED_L=$L timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline \
        > gpurun_out/sweep4_${lib}_L$L.json 2> gpurun_out/sweep4_${lib}_L$L.err
  done
done
unset MGVIB200_LIB
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --iters-per-step 5"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:k_field_pass -s 48 -c 3 -o gpurun_out/prof_r1b $CMD > gpurun_out/ncu_full.log 2>&1
tail -n 3 gpurun_out/smoke.log gpurun_out/pytest_gpu.log
