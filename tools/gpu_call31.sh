#!/bin/bash
# round-2 call 31: sample groups on 1-D fields (pair-aligned split): bitwise test, C2 timing off / on, 1-D parity tests
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "two_sample_groups or c2 or long_1d or metric_1d or golden" > gpurun_out/r2_pytest_groups1d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_groups1d.log)
tail -n 4 gpurun_out/r2_pytest_groups1d.log | cut -c1-300
Q="--steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --no-configs --workload c2 --iters-per-step 200"
for rep in a b; do
for gm in 1 -1; do
  MGVIB200_GROUPS=$gm timeout 300 python bench.py $Q > gpurun_out/r2_c2_g${gm}$rep.json 2> gpurun_out/r2_c2_g${gm}$rep.err
  python - gpurun_out/r2_c2_g${gm}$rep.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value %.1f ms/iter %.4f" % (d["value"], d["cg_path"]["lockstep_iteration_ms"]), d["clocks"]["sm_mhz"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
done
tail -n 3 gpurun_out/r2_c2_g-1a.err | cut -c1-300
