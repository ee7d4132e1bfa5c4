#!/bin/bash
# round-2 call 7 (2 GPUs): rank-count independence (bitwise 1 vs 2 GPUs) and a 2-GPU bench line
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py > gpurun_out/r2_multigpu_check_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_multigpu_check_2gpu.log
grep -E "dims|PASS|FAIL|rc=|Error|error" gpurun_out/r2_multigpu_check_2gpu.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_bench_2gpu.json")); print("2gpu value %.1f ms/iter %.4f e2e %.1f" % (d["value"], d["cg_path"]["lockstep_iteration_ms"], d["e2e"]["value"])); print(json.dumps(d["configs"])[:1500]); print(d["mgvi_step"]["seconds"], d["mgvi_step"]["newton_ms"])
except Exception as e: print("ERR", e)
PY
tail -n 5 gpurun_out/r2_bench_2gpu.err | cut -c1-300
