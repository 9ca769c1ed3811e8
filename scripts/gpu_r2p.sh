#!/bin/bash
# N=1: operator pass A/B, spectral + FIR GPU tests
set -x
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.txt 2>&1; echo "smoke rc=$?" > gpurun_out/r2p_summary.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "spectral or fir or fused or canonical or c5 or preset" > gpurun_out/r2p_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/r2p_summary.txt
MS_SPEC_PASS=1 timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 0 --cpu-sample 0 > gpurun_out/r2p_bench_pass1.json 2> gpurun_out/r2p_bench_pass1.err; echo "pass1 rc=$?" >> gpurun_out/r2p_summary.txt
MS_SPEC_PASS=0 timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 0 --cpu-sample 0 > gpurun_out/r2p_bench_pass0.json 2> gpurun_out/r2p_bench_pass0.err; echo "pass0 rc=$?" >> gpurun_out/r2p_summary.txt
tail -3 gpurun_out/r2p_tests.txt; cat gpurun_out/r2p_summary.txt
