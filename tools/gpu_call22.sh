#!/bin/bash
# round-2 call 22: compile-time-length Bluestein kernels: parity + 1000 x 1000 timing
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -k "any_length or field_ops or golden" > gpurun_out/r2_pytest_blue2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_blue2.log)
tail -n 3 gpurun_out/r2_pytest_blue2.log | cut -c1-300
Q="--steps 2 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --no-configs --workload np2 --iters-per-step 10"
timeout 300 python bench.py $Q > gpurun_out/r2_np2_blue2.json 2> gpurun_out/r2_np2_blue2.err
python - gpurun_out/r2_np2_blue2.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value %.1f ms/iter %.4f" % (d["value"], d["cg_path"]["lockstep_iteration_ms"]), [(k["tag"], round(k["avg_ms"],4), k["launches"]) for k in d["kernels"][:8]])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
