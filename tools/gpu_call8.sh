#!/bin/bash
# round-2 call 8: persistent cooperative CG kernel for C2 (parity + timing on/off), MINB=4 variant for the row passes
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "c2 or long_1d or golden" > gpurun_out/r2_pytest_persist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_persist.log)
tail -n 6 gpurun_out/r2_pytest_persist.log | cut -c1-300
Q="--steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --no-configs"
for ps in 0 1; do
  MGVIB200_PERSIST=$ps timeout 300 python bench.py $Q --workload c2 --iters-per-step 200 > gpurun_out/r2_c2_persist$ps.json 2> gpurun_out/r2_c2_persist$ps.err
done
for lib in default mb4; do
  L=""; if [ $lib = mb4 ]; then L=build/variants/libmgvib200_mb4.so; fi
  MGVIB200_LIB=$L timeout 300 python bench.py $Q > gpurun_out/r2_c3_$lib.json 2> gpurun_out/r2_c3_$lib.err
  MGVIB200_LIB=$L timeout 300 python bench.py $Q --n-residuals 4 > gpurun_out/r2_c3n4_$lib.json 2> gpurun_out/r2_c3n4_$lib.err
  MGVIB200_LIB=$L timeout 300 python bench.py $Q --workload c5 --n-residuals 8 --iters-per-step 10 > gpurun_out/r2_c5n8_$lib.json 2> gpurun_out/r2_c5n8_$lib.err
done
for f in r2_c2_persist0 r2_c2_persist1 r2_c3_default r2_c3_mb4 r2_c3n4_default r2_c3n4_mb4 r2_c5n8_default r2_c5n8_mb4; do python - gpurun_out/$f.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value %.1f ms/iter %.4f" % (d["value"], d["cg_path"]["lockstep_iteration_ms"]), [(k["kernel"][:12], round(k["avg_ms"],4)) for k in d["kernels"][:6]])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -n 3 gpurun_out/r2_c2_persist1.err | cut -c1-300
