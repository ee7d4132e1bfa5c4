#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
for lib in default mb4; do
  for L in 2 4; do
    for PF in 0 1 2; do
      if [ "$lib" = "default" ]; then unset MGVIB200_LIB; else export MGVIB200_LIB=$PWD/build/variants/libmgvib200_$lib.so; fi
      MGVIB200_STRIDED_L=$L MGVIB200_PREFETCH_WAVES=$PF timeout 300 python bench.py --steps 2 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline \
          > gpurun_out/sweep5_${lib}_L${L}_pf$PF.json 2> gpurun_out/sweep5_${lib}_L${L}_pf$PF.err
    done
  done
done
unset MGVIB200_LIB
tail -n 3 gpurun_out/pytest_gpu.log
