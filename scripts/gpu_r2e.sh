#!/bin/bash
# round 2, call E (2 GPUs): full GPU test-suite, N=1 bench with e2e (native planner), N=2 bench variants
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2e_summary.txt
tail -5 gpurun_out/r2e_pytest.log
MS_TRACE=1 timeout 900 python bench.py --steps 5 --warmup 3 --e2e-steps 3 --cpu-sample 0 > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; echo "bench n1 rc=$?" >> gpurun_out/r2e_summary.txt
for v in "" "--no-graph --slices 1" "--slices 2" "--slices 8"; do
  tag=$(echo "$v" | tr -d ' -')
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --e2e-steps 2 --cpu-sample 0 $v > gpurun_out/r2e_bench_n2_$tag.json 2> gpurun_out/r2e_bench_n2_$tag.err; echo "bench n2 [$v] rc=$?" >> gpurun_out/r2e_summary.txt
done
cat gpurun_out/r2e_summary.txt
