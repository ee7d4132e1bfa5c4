#!/bin/bash
# round-2 call 2: the whole GPU parity suite (no -x) on the committed build
mkdir -p gpurun_out
(timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu2.log)
tail -n 40 gpurun_out/r2_pytest_gpu2.log
