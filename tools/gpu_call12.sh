#!/bin/bash
# round-2 call 12: C2 block-size sweep of the four-step kernels and of cg_update
mkdir -p gpurun_out
Q="--steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --no-configs --workload c2 --iters-per-step 200"
for bt in 256 128 64; do
 for vp in 4096 2048 1772; do
  MGVIB200_BIG_THREADS=$bt MGVIB200_VEC_PER_BLOCK=$vp timeout 300 python bench.py $Q > gpurun_out/r2_c2_bt${bt}_vp$vp.json 2> gpurun_out/r2_c2_bt${bt}_vp$vp.err
  python - gpurun_out/r2_c2_bt${bt}_vp$vp.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value %.1f ms/iter %.4f" % (d["value"], d["cg_path"]["lockstep_iteration_ms"]), [(k["tag"], round(k["avg_ms"],4)) for k in d["kernels"][:6]])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
 done
done
