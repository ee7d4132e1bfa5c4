#!/bin/bash
# round-2 call 37: the default bench line on the last build (configs block with 1 + 8k iteration counts)
mkdir -p gpurun_out
(time timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err) 2> gpurun_out/bench_final.time
tail -c 300 gpurun_out/bench_final.err; head -2 gpurun_out/bench_final.time
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_final.json"))
print(d["value"], d["cg_path"]["lockstep_iteration_ms"], d["e2e"]["value"], d["clocks"]["sm_mhz"])
print({k:(v.get("lockstep_iteration_ms") or v.get("mgvi_step_ms") or v.get("syrk_ms") or v.get("seconds")) for k,v in d["configs"].items()})
print({k:v for k,v in d["configs"]["c5"].items() if k!="workload"})
PY
