#!/bin/bash
# round-2 call 19: GEMM-based blocked Cholesky (inverted diagonal blocks, K = 256 trailing updates): dense parity + timing
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "dense or matrix_inversion or per_datum" > gpurun_out/r2_pytest_chol.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_chol.log)
tail -n 4 gpurun_out/r2_pytest_chol.log | cut -c1-300
timeout 300 python tools/bench_dense.py > gpurun_out/r2_dense_chol2.json 2> gpurun_out/r2_dense_chol2.err; cut -c1-900 gpurun_out/r2_dense_chol2.json; tail -n 3 gpurun_out/r2_dense_chol2.err
