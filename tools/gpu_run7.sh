#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/smi2.txt
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > gpurun_out/multigpu_check.log 2>&1; echo "rc=$?" >> gpurun_out/multigpu_check.log)
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 1 > gpurun_out/bench_c3_2gpu.json 2> gpurun_out/bench_c3_2gpu.err; echo "rc=$?" >> gpurun_out/bench_c3_2gpu.err)
(timeout 900 python -m pytest tests -m gpu -q -k "c2_1d or long_1d" > gpurun_out/pytest_c2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_c2.log)
(timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 --e2e-steps 2 --iters-per-step 100 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "rc=$?" >> gpurun_out/bench_c2.err)
tail -n 4 gpurun_out/multigpu_check.log gpurun_out/bench_c3_2gpu.err gpurun_out/pytest_c2.log gpurun_out/bench_c2.err
