#!/bin/bash
# round-2 call 16: deferred solution update (x += alpha p inside the fused axis-0 pass): parity + timing on/off
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -k "field_ops or golden or many_residuals or default_tolerance or full_size_c3 or large_2d" > gpurun_out/r2_pytest_defer.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_defer.log)
tail -n 3 gpurun_out/r2_pytest_defer.log | cut -c1-300
Q="--steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --no-configs"
for df in 0 1 2; do
  MGVIB200_DEFER_X=$df timeout 300 python bench.py $Q > gpurun_out/r2_c3_defer$df.json 2> gpurun_out/r2_c3_defer$df.err
  MGVIB200_DEFER_X=$df timeout 300 python bench.py $Q --n-residuals 4 > gpurun_out/r2_c3n4_defer$df.json 2> gpurun_out/r2_c3n4_defer$df.err
  MGVIB200_DEFER_X=$df timeout 300 python bench.py $Q --workload c5 --n-residuals 8 --iters-per-step 10 > gpurun_out/r2_c5n8_defer$df.json 2> gpurun_out/r2_c5n8_defer$df.err
done
for f in r2_c3_defer0 r2_c3_defer1 r2_c3_defer2 r2_c3n4_defer0 r2_c3n4_defer1 r2_c3n4_defer2 r2_c5n8_defer0 r2_c5n8_defer1 r2_c5n8_defer2; do python - gpurun_out/$f.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value %.1f ms/iter %.4f" % (d["value"], d["cg_path"]["lockstep_iteration_ms"]), [(k["tag"], round(k["avg_ms"],4)) for k in d["kernels"][:6]], d["clocks"]["sm_mhz"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -n 3 gpurun_out/r2_c3_defer1.err | cut -c1-300
