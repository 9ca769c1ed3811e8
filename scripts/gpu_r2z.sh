#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.txt 2>&1; echo "smoke rc=$?" > gpurun_out/r2z_summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2z_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/r2z_summary.txt
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 0 --cpu-sample 0 > gpurun_out/r2z_$tag.json 2> gpurun_out/r2z_$tag.err; echo "$tag rc=$?" >> gpurun_out/r2z_summary.txt; }
run base FOO=1
run w1 MS_PLAN_W=1.1,1.0,1.4,2.4,0.02
run w2 MS_PLAN_W=1.1,1.0,1.2,3.0,0.02
tail -5 gpurun_out/r2z_tests.txt; cat gpurun_out/r2z_summary.txt
