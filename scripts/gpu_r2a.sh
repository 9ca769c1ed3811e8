#!/bin/bash
# round 2, call A: parity of the fused FIR path in every launch mode, then the bench per mode
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
for m in 8 1 4 16 0; do
  MS_FIR_MODE=$m timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke_m$m.log 2>&1; echo "smoke mode $m rc=$?" >> gpurun_out/r2a_summary.txt
done
for m in 8 1 4 16 0; do
  MS_FIR_MODE=$m timeout 900 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --cpu-sample 0 > gpurun_out/r2a_bench_m$m.json 2> gpurun_out/r2a_bench_m$m.err; echo "bench mode $m rc=$?" >> gpurun_out/r2a_summary.txt
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_summary.txt
tail -5 gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_summary.txt
