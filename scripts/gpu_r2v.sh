#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.txt 2>&1; echo "smoke rc=$?" > gpurun_out/r2v_summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2v_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/r2v_summary.txt
for cl in 1 4; do MS_SYNTH_CLUSTER=$cl timeout 300 python bench.py --renders 512 --steps 10 --warmup 3 --e2e-steps 0 --cpu-sample 0 > gpurun_out/r2v_b512_cl$cl.json 2> gpurun_out/r2v_b512_cl$cl.err; echo "b512 cl$cl rc=$?" >> gpurun_out/r2v_summary.txt; done
timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 0 --cpu-sample 0 > gpurun_out/r2v_n1.json 2> gpurun_out/r2v_n1.err; echo "n1 rc=$?" >> gpurun_out/r2v_summary.txt
tail -5 gpurun_out/r2v_tests.txt; cat gpurun_out/r2v_summary.txt
