#!/bin/bash
# round-2 call 3: new model family + MatrixInversion on fields + per-datum dense F, then the whole GPU suite
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "hyper or coal or matrix_inversion or per_datum or mgvi_sample or polyfit" > gpurun_out/r2_pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_new.log)
tail -n 30 gpurun_out/r2_pytest_new.log
(timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu3.log)
tail -n 5 gpurun_out/r2_pytest_gpu3.log
