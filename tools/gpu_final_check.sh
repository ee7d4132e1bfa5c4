#!/bin/bash
# round-end style check: GPU parity suite, smoke(), the default bench line and the reference arm
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log)
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log)
(time timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err) 2> gpurun_out/bench_final.time
(time timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err) 2> gpurun_out/bench_reference.time
tail -n 2 gpurun_out/pytest_gpu.log gpurun_out/smoke.log
