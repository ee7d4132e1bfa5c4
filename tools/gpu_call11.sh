#!/bin/bash
# round-2 call 11: loads-before-stores in the four-step CG prologue / metric epilogue, unrolled cg_update: C2 + C3 timing, parity
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "c2 or long_1d or field_ops or golden" > gpurun_out/r2_pytest_c11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_c11.log)
tail -n 3 gpurun_out/r2_pytest_c11.log | cut -c1-300
Q="--steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --no-configs"
for ps in 0 1; do
  MGVIB200_PERSIST=$ps timeout 300 python bench.py $Q --workload c2 --iters-per-step 200 > gpurun_out/r2_c2_c11_p$ps.json 2> gpurun_out/r2_c2_c11_p$ps.err
done
timeout 300 python bench.py $Q > gpurun_out/r2_c3_c11.json 2> gpurun_out/r2_c3_c11.err
timeout 300 python bench.py $Q --n-residuals 4 > gpurun_out/r2_c3n4_c11.json 2> gpurun_out/r2_c3n4_c11.err
for f in r2_c2_c11_p0 r2_c2_c11_p1 r2_c3_c11 r2_c3n4_c11; do python - gpurun_out/$f.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value %.1f ms/iter %.4f" % (d["value"], d["cg_path"]["lockstep_iteration_ms"]), [(k["tag"], round(k["avg_ms"],4)) for k in d["kernels"][:6]])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
