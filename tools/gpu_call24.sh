#!/bin/bash
# round-2 call 24: conflict-free first exchange of the first pass: parity + C3 / C5 timing
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -k "field_ops or golden or full_size or large_2d or deferred" > gpurun_out/r2_pytest_swz.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_swz.log)
tail -n 3 gpurun_out/r2_pytest_swz.log | cut -c1-200
Q="--steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --no-configs"
for rep in a b; do
  timeout 300 python bench.py $Q > gpurun_out/r2_c3_swz$rep.json 2> gpurun_out/r2_c3_swz$rep.err
  python - gpurun_out/r2_c3_swz$rep.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value %.1f ms/iter %.4f" % (d["value"], d["cg_path"]["lockstep_iteration_ms"]), [(k["tag"], round(k["avg_ms"],4)) for k in d["kernels"][:4]], d["clocks"]["sm_mhz"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
timeout 300 python bench.py $Q --workload c5 --n-residuals 8 --iters-per-step 10 > gpurun_out/r2_c5n8_swz.json 2> gpurun_out/r2_c5n8_swz.err
python - gpurun_out/r2_c5n8_swz.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], "value %.1f ms/iter %.4f" % (d["value"], d["cg_path"]["lockstep_iteration_ms"]), [(k["tag"], round(k["avg_ms"],4)) for k in d["kernels"][:6]], d["clocks"]["sm_mhz"])
PY
