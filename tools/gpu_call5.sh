#!/bin/bash
# round-2 call 5: paired-arrangement transforms A/B, dense GEMM check, parity of the new kernels
mkdir -p gpurun_out
Q="--steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --no-configs"
for pf in 0 1; do
  MGVIB200_PF=$pf timeout 300 python bench.py $Q > gpurun_out/r2_c3_pf$pf.json 2> gpurun_out/r2_c3_pf$pf.err
  MGVIB200_PF=$pf timeout 300 python bench.py $Q --n-residuals 4 > gpurun_out/r2_c3n4_pf$pf.json 2> gpurun_out/r2_c3n4_pf$pf.err
  MGVIB200_PF=$pf timeout 300 python bench.py $Q --workload c5 --n-residuals 8 --iters-per-step 10 > gpurun_out/r2_c5n8_pf$pf.json 2> gpurun_out/r2_c5n8_pf$pf.err
done
for f in gpurun_out/r2_c3_pf0.json gpurun_out/r2_c3_pf1.json gpurun_out/r2_c3n4_pf0.json gpurun_out/r2_c3n4_pf1.json gpurun_out/r2_c5n8_pf0.json gpurun_out/r2_c5n8_pf1.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value %.1f ms/iter %.4f" % (d["value"], d["cg_path"]["lockstep_iteration_ms"]), [(k["kernel"][:12], round(k["avg_ms"],4)) for k in d["kernels"][:6]])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
timeout 300 python tools/bench_dense.py > gpurun_out/r2_dense.json 2> gpurun_out/r2_dense.err; cut -c1-600 gpurun_out/r2_dense.json
(timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu5.log)
tail -n 6 gpurun_out/r2_pytest_gpu5.log | cut -c1-300
