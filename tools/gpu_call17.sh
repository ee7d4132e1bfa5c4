#!/bin/bash
# round-2 call 17: deferred update defaults, bitwise GPU test, C3 timing on/off
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -k "deferred or full_size_c3 or many_residuals" > gpurun_out/r2_pytest_defer2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_defer2.log)
tail -n 3 gpurun_out/r2_pytest_defer2.log | cut -c1-300
Q="--steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --no-configs"
for rep in a b; do
for df in 0 1; do
  MGVIB200_DEFER_X=$df timeout 300 python bench.py $Q > gpurun_out/r2_c3_dfr${df}$rep.json 2> gpurun_out/r2_c3_dfr${df}$rep.err
  python - gpurun_out/r2_c3_dfr${df}$rep.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value %.1f ms/iter %.4f" % (d["value"], d["cg_path"]["lockstep_iteration_ms"]), [(k["tag"], round(k["avg_ms"],4)) for k in d["kernels"][:6]], d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
done
