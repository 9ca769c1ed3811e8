#!/bin/bash
# round 2, call B: Hermitian phase 2 + post recompute; parity then bench (phases vs cluster-8)
set -x
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?" > gpurun_out/r2b_summary.txt
MS_FIR_MODE=8 timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke_m8.log 2>&1; echo "smoke m8 rc=$?" >> gpurun_out/r2b_summary.txt
for m in 1 8; do
  MS_FIR_MODE=$m timeout 900 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --cpu-sample 0 > gpurun_out/r2b_bench_m$m.json 2> gpurun_out/r2b_bench_m$m.err; echo "bench mode $m rc=$?" >> gpurun_out/r2b_summary.txt
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_summary.txt
tail -5 gpurun_out/r2b_pytest.log
cat gpurun_out/r2b_summary.txt
