#!/bin/bash
# round-2 call 21: ncu launch list + --set full of the CG kernels on the final build (deferred solution update on)
mkdir -p gpurun_out
export MGVIB200_GRAPH=0
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --no-configs --iters-per-step 5"
$CMD > gpurun_out/r2b_ncu_plain.json 2> gpurun_out/r2b_ncu_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_c3.csv $CMD > gpurun_out/r2b_ncu_launches.log 2>&1
$CMD > gpurun_out/r2b_ncu_plain2.json 2> gpurun_out/r2b_ncu_plain2.err && \
ncu --set full --clock-control none --import-source on -k regex:"k_field_pass|k_vec" -s 60 -c 4 -o /tmp/prof_r2b $CMD > gpurun_out/r2b_ncu_full.log 2>&1
ncu -i /tmp/prof_r2b.ncu-rep --page raw --csv > gpurun_out/r2b_prof_raw.csv 2>/dev/null
tail -n 2 gpurun_out/r2b_ncu_full.log; ls -la gpurun_out/r2b_prof_raw.csv
unset MGVIB200_GRAPH
python bench.py --steps 5 --warmup 3 --no-e2e --no-full-step --no-cpu-baseline --no-configs > gpurun_out/r2b_bench_c3_short.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2b_bench_c3_short.json")); print(d["value"], d["cg_path"]["lockstep_iteration_ms"], d["roofline"]["kernel"], d["roofline"]["frac"], [(k["kernel"][:16], round(k["avg_ms"],4)) for k in d["kernels"][:5]], d["clocks"])
PY
