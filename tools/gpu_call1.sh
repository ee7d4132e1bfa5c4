#!/bin/bash
# round-2 call 1: smoke, GPU parity suite, default bench, graph on/off and twiddle-squaring comparisons
mkdir -p gpurun_out
(timeout 400 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log)
Q="--steps 3 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline"
for g in 1 0; do
  MGVIB200_GRAPH=$g timeout 300 python bench.py $Q > gpurun_out/r2_c3_graph$g.json 2> gpurun_out/r2_c3_graph$g.err
  MGVIB200_GRAPH=$g timeout 300 python bench.py $Q --workload c2 --iters-per-step 200 > gpurun_out/r2_c2_graph$g.json 2> gpurun_out/r2_c2_graph$g.err
done
MGVIB200_LIB=$PWD/build/variants/libmgvib200_twsq.so timeout 300 python bench.py $Q > gpurun_out/r2_c3_twsq.json 2> gpurun_out/r2_c3_twsq.err
timeout 300 python bench.py $Q > gpurun_out/r2_c3_base2.json 2> gpurun_out/r2_c3_base2.err
(timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log)
(time timeout 600 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err) 2> gpurun_out/r2_bench_default.time
tail -n 3 gpurun_out/r2_smoke.log gpurun_out/r2_pytest_gpu.log
for f in gpurun_out/r2_c3_graph1.json gpurun_out/r2_c3_graph0.json gpurun_out/r2_c2_graph1.json gpurun_out/r2_c2_graph0.json gpurun_out/r2_c3_twsq.json gpurun_out/r2_c3_base2.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "value %.1f ms/iter %.4f" % (d["value"], d["cg_path"]["lockstep_iteration_ms"]), [(k["kernel"][:12], round(k["avg_ms"],4)) for k in d["kernels"][:5]])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
