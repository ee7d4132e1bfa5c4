#!/bin/bash
# round-2 call 29: two concurrent sample groups: bitwise test, whole GPU suite, C3 timing at 4 / 8 / 32 samples with groups off / on
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "two_sample_groups or deferred" > gpurun_out/r2_pytest_groups.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_groups.log)
tail -n 4 gpurun_out/r2_pytest_groups.log | cut -c1-300
Q="--steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --no-configs"
for n in 4 8 32; do
for gm in 1 2; do
  MGVIB200_GROUPS=$gm timeout 300 python bench.py $Q --n-residuals $n > gpurun_out/r2_c3n${n}_g$gm.json 2> gpurun_out/r2_c3n${n}_g$gm.err
  python - gpurun_out/r2_c3n${n}_g$gm.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value %.1f ms/iter %.4f" % (d["value"], d["cg_path"]["lockstep_iteration_ms"]), d["clocks"]["sm_mhz"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
done
tail -n 3 gpurun_out/r2_c3n4_g2.err | cut -c1-300
(timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu7.log)
tail -n 3 gpurun_out/r2_pytest_gpu7.log | cut -c1-300
