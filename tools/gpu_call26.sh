#!/bin/bash
# round-2 call 26: stream-K schedule of the J'FJ rank-m update: dense parity + SYRK timing on/off
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "dense or matrix_inversion or per_datum" > gpurun_out/r2_pytest_sk.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_sk.log)
tail -n 3 gpurun_out/r2_pytest_sk.log | cut -c1-200
for sk in 0 1; do
  MGVIB200_STREAMK=$sk timeout 300 python tools/bench_dense.py > gpurun_out/r2_dense_sk$sk.json 2> gpurun_out/r2_dense_sk$sk.err
  python - gpurun_out/r2_dense_sk$sk.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "syrk %.3f ms (%.2f TF/s half) chol %.2f ms solve %.2f ms step %.4f s mnlp %.6f" % (d["syrk_ms"], d["syrk_tflops_symmetric_half"], d["cholesky_ms"], d["sample_ms_after_factor"], d["mgvi_step_s"], d["mnlp"]))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
  tail -n 2 gpurun_out/r2_dense_sk$sk.err | cut -c1-300
done
