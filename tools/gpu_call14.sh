#!/bin/bash
# round-2 call 14: full GPU suite on the current build; 1000 x 1000 Bluestein vs direct pass timing
mkdir -p gpurun_out
Q="--steps 2 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --no-configs --workload np2 --iters-per-step 10"
timeout 300 python bench.py $Q > gpurun_out/r2_np2_blue.json 2> gpurun_out/r2_np2_blue.err
MGVIB200_BLUE_MIN=100000 timeout 600 python bench.py $Q > gpurun_out/r2_np2_direct.json 2> gpurun_out/r2_np2_direct.err
for f in r2_np2_blue r2_np2_direct; do python - gpurun_out/$f.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value %.1f ms/iter %.4f" % (d["value"], d["cg_path"]["lockstep_iteration_ms"]), [(k["tag"], round(k["avg_ms"],4), k["launches"]) for k in d["kernels"][:8]])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
(timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu6.log)
tail -n 6 gpurun_out/r2_pytest_gpu6.log | cut -c1-300
