#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_dense.py > gpurun_out/dense16.json 2> gpurun_out/dense16.err
(timeout 900 python -m pytest tests -m gpu -q -x -k "dense or degenerate or empty or polyfit" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
tail -n 3 gpurun_out/pytest_gpu.log
