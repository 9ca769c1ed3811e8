#!/bin/bash
# round 2, call H: warp-local Bluestein columns (B1 = 256) A/B, parity of the spectral tests, e2e with the finer slice ramp
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fft_lengths or spectral_ops or c5_members or golden" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2h_summary.txt
tail -3 gpurun_out/r2h_pytest.log
MS_SB_WARP=0 timeout 600 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --cpu-sample 0 > gpurun_out/r2h_bench_old.json 2> gpurun_out/r2h_bench_old.err; echo "bench old rc=$?" >> gpurun_out/r2h_summary.txt
MS_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --e2e-steps 4 --cpu-sample 0 > gpurun_out/r2h_bench_new.json 2> gpurun_out/r2h_bench_new.err; echo "bench new rc=$?" >> gpurun_out/r2h_summary.txt
cat gpurun_out/r2h_summary.txt
