#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
(timeout 600 python tools/bench_dense.py > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err; echo "rc=$?" >> gpurun_out/bench_dense.err)
(timeout 900 python bench.py --steps 5 --warmup 3 --e2e-steps 1 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "rc=$?" >> gpurun_out/bench_c3.err)
(timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?" >> gpurun_out/bench_ref.err)
DENSE_M=8192 DENSE_K=2048 python tools/bench_dense.py > gpurun_out/ncu_dense_plain.json 2> gpurun_out/ncu_dense_plain.err && \
DENSE_M=8192 DENSE_K=2048 ncu --set full --clock-control none --import-source on -k regex:k_dmma_gemm -c 1 -o gpurun_out/prof_r1_dense python tools/bench_dense.py > gpurun_out/ncu_dense.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/bench_dense.err gpurun_out/bench_c3.err
