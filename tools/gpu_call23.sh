#!/bin/bash
# round-2 call 23: loads in flight per thread in the deferred-update prologue (2 / 4 / 8 quadruples)
mkdir -p gpurun_out
Q="--steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --no-configs"
for rep in a b; do
for v in default du2 du8; do
  L=""; if [ $v != default ]; then L=build/variants/libmgvib200_$v.so; fi
  MGVIB200_LIB=$L timeout 300 python bench.py $Q > gpurun_out/r2_c3_$v$rep.json 2> gpurun_out/r2_c3_$v$rep.err
  python - gpurun_out/r2_c3_$v$rep.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value %.1f ms/iter %.4f" % (d["value"], d["cg_path"]["lockstep_iteration_ms"]), [(k["tag"], round(k["avg_ms"],4)) for k in d["kernels"][:4]], d["clocks"]["sm_mhz"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
done
