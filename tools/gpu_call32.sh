#!/bin/bash
# round-2 call 32: sample groups at C5 (8 samples per GPU) off / on
mkdir -p gpurun_out
Q="--steps 2 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --no-configs --workload c5 --n-residuals 8 --iters-per-step 10"
for rep in a b; do
for gm in 1 -1; do
  MGVIB200_GROUPS=$gm timeout 300 python bench.py $Q > gpurun_out/r2_c5n8_g${gm}$rep.json 2> gpurun_out/r2_c5n8_g${gm}$rep.err
  python - gpurun_out/r2_c5n8_g${gm}$rep.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value %.1f ms/iter %.4f" % (d["value"], d["cg_path"]["lockstep_iteration_ms"]), d["clocks"]["sm_mhz"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
done
