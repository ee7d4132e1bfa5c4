#!/bin/bash
# round-2 call 9: A' column kernel with 4 neighbouring columns + mirrors per unit (C2), parity + timing
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "c2 or long_1d" > gpurun_out/r2_pytest_ap8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_ap8.log)
tail -n 4 gpurun_out/r2_pytest_ap8.log | cut -c1-300
Q="--steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --no-configs"
for lg in 1 2 3; do
 for ps in 0 1; do
  MGVIB200_BIG_AP_LGL=$lg MGVIB200_PERSIST=$ps timeout 300 python bench.py $Q --workload c2 --iters-per-step 200 > gpurun_out/r2_c2_ap${lg}_p$ps.json 2> gpurun_out/r2_c2_ap${lg}_p$ps.err
  python - gpurun_out/r2_c2_ap${lg}_p$ps.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value %.1f ms/iter %.4f" % (d["value"], d["cg_path"]["lockstep_iteration_ms"]), [(k["tag"], round(k["avg_ms"],4)) for k in d["kernels"][:6]])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
 done
done
