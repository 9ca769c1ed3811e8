#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1; echo "smoke rc=$?" > gpurun_out/r2d_summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --cpu-sample 0 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?" >> gpurun_out/r2d_summary.txt
cat gpurun_out/r2d_summary.txt
