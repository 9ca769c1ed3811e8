#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3a_smoke.txt 2>&1; echo "smoke rc=$?" > gpurun_out/r3a_summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r3a_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/r3a_summary.txt
timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 0 --cpu-sample 0 > gpurun_out/r3a_n1.json 2> gpurun_out/r3a_n1.err; echo "n1 rc=$?" >> gpurun_out/r3a_summary.txt
tail -5 gpurun_out/r3a_tests.txt; cat gpurun_out/r3a_summary.txt
