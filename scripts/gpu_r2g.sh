#!/bin/bash
# round 2, call G (2 GPUs): pipelined gather at N=2, e2e with the ramped-down slice schedule at N=1
set -x
mkdir -p gpurun_out
MS_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --e2e-steps 4 --cpu-sample 0 > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; echo "bench n1 rc=$?" > gpurun_out/r2g_summary.txt
for v in "" "--no-pipeline"; do
  tag=$(echo "$v" | tr -d ' -')
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --e2e-steps 2 --cpu-sample 0 $v > gpurun_out/r2g_bench_n2_$tag.json 2> gpurun_out/r2g_bench_n2_$tag.err; echo "bench n2 [$v] rc=$?" >> gpurun_out/r2g_summary.txt
done
cat gpurun_out/r2g_summary.txt
