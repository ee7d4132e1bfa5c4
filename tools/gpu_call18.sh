#!/bin/bash
# round-2 call 18 (8 GPUs): rank-count independence at 8 ranks, the default bench line at N = 8 (incl. the C5 step with its all-gathers)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/multigpu_check.py > gpurun_out/r2_multigpu_check_8gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_multigpu_check_8gpu.log
grep -E "dims|PASS|FAIL|rc=|Error|error" gpurun_out/r2_multigpu_check_8gpu.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_bench_8gpu.json")); print("8gpu value %.1f ms/iter %.4f e2e %.1f" % (d["value"], d["cg_path"]["lockstep_iteration_ms"], d["e2e"]["value"])); print(json.dumps(d["configs"])[:1800]); print(d["mgvi_step"]["seconds"], d["mgvi_step"]["newton_ms"])
except Exception as e: print("ERR", e)
PY
tail -n 5 gpurun_out/r2_bench_8gpu.err | cut -c1-300
