#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > gpurun_out/multigpu_check_$N.log 2>&1; echo "rc=$?" >> gpurun_out/multigpu_check_$N.log)
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --e2e-steps 1 > gpurun_out/bench_c3_${N}gpu.json 2> gpurun_out/bench_c3_${N}gpu.err; echo "rc=$?" >> gpurun_out/bench_c3_${N}gpu.err)
grep -E "dims|PASS|FAIL|rc=" gpurun_out/multigpu_check_$N.log | cut -c1-400; tail -n 2 gpurun_out/bench_c3_${N}gpu.err
