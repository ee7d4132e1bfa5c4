#!/bin/bash
# round-2 call 25 (4 GPUs): scaling points N = 2 and N = 4 on the final build (kernel-only + e2e legs, no configs block)
mkdir -p gpurun_out
for n in 2 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 5 --warmup 3 --no-configs > gpurun_out/r2b_bench_${n}gpu.json 2> gpurun_out/r2b_bench_${n}gpu.err; echo "bench $n rc=$?"
  python - $n <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open("gpurun_out/r2b_bench_%sgpu.json"%n)); print(n, "gpu value %.1f ms/iter %.4f e2e %.1f step %.2fs newton %.1f ms" % (d["value"], d["cg_path"]["lockstep_iteration_ms"], d["e2e"]["value"], d["mgvi_step"]["seconds"], d["mgvi_step"]["newton_ms"]), d["clocks"]["sm_mhz"])
except Exception as e: print("ERR", e)
PY
done
