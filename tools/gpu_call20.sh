#!/bin/bash
# round-2 call 20: ncu of the small DMMA GEMM launches and potrf inside the blocked Cholesky
mkdir -p gpurun_out
export DENSE_M=4096 DENSE_K=4096
timeout 300 python tools/bench_dense.py > gpurun_out/r2_dense_small.json 2> gpurun_out/r2_dense_small.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_dense.csv python tools/bench_dense.py > gpurun_out/r2_ncu_dense_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_dmma_gemm|k_potrf" -s 10 -c 12 -o /tmp/prof_dense python tools/bench_dense.py > gpurun_out/r2_ncu_dense_full.log 2>&1
ncu -i /tmp/prof_dense.ncu-rep --page raw --csv > gpurun_out/r2_prof_dense_raw.csv 2>/dev/null
tail -n 2 gpurun_out/r2_ncu_dense_full.log; ls -la gpurun_out/r2_prof_dense_raw.csv gpurun_out/r2_launches_dense.csv
