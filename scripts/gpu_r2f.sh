#!/bin/bash
# round 2, call F: end-to-end variants at N=1 (slice size, planner threads)
set -x
mkdir -p gpurun_out
i=0
for v in "--chunk 512" "--chunk 1024" "--chunk 2048"; do
  for th in 2 4; do
    i=$((i+1))
    MS_PLAN_THREADS=$th MS_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --e2e-steps 3 --cpu-sample 0 $v > gpurun_out/r2f_bench_$i.json 2> gpurun_out/r2f_bench_$i.err; echo "bench $i [$v threads $th] rc=$?" >> gpurun_out/r2f_summary.txt
  done
done
cat gpurun_out/r2f_summary.txt
