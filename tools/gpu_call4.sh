#!/bin/bash
# round-2 call 4: new-feature tests, PDL on/off, 4-sample (= per-GPU share at N=8) tuning runs, default bench, reference arm
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "hyper or coal or matrix_inversion or per_datum or mgvi_sample or polyfit" > gpurun_out/r2_pytest_new2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_new2.log)
tail -n 12 gpurun_out/r2_pytest_new2.log | cut -c1-400
Q="--steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --no-configs"
for pdl in 0 1; do
  MGVIB200_PDL=$pdl timeout 300 python bench.py $Q > gpurun_out/r2_c3_pdl$pdl.json 2> gpurun_out/r2_c3_pdl$pdl.err
  MGVIB200_PDL=$pdl timeout 300 python bench.py $Q --n-residuals 4 > gpurun_out/r2_c3n4_pdl$pdl.json 2> gpurun_out/r2_c3n4_pdl$pdl.err
  MGVIB200_PDL=$pdl timeout 300 python bench.py $Q --workload c2 --iters-per-step 200 > gpurun_out/r2_c2_pdl$pdl.json 2> gpurun_out/r2_c2_pdl$pdl.err
done
MGVIB200_PDL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "field_ops or c2 or golden or many_residuals" > gpurun_out/r2_pytest_pdl.log 2>&1; tail -n 2 gpurun_out/r2_pytest_pdl.log
for f in gpurun_out/r2_c3_pdl0.json gpurun_out/r2_c3_pdl1.json gpurun_out/r2_c3n4_pdl0.json gpurun_out/r2_c3n4_pdl1.json gpurun_out/r2_c2_pdl0.json gpurun_out/r2_c2_pdl1.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value %.1f ms/iter %.4f" % (d["value"], d["cg_path"]["lockstep_iteration_ms"]), [(k["kernel"][:12], round(k["avg_ms"],4)) for k in d["kernels"][:5]])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
(time timeout 900 python bench.py > gpurun_out/r2_bench_default2.json 2> gpurun_out/r2_bench_default2.err) 2> gpurun_out/r2_bench_default2.time
tail -c 600 gpurun_out/r2_bench_default2.err; cat gpurun_out/r2_bench_default2.time
(time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference2.json 2> gpurun_out/r2_bench_reference2.err) 2> gpurun_out/r2_bench_reference2.time
cat gpurun_out/r2_bench_reference2.json | cut -c1-700; cat gpurun_out/r2_bench_reference2.time; nproc
