#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "planner or batch or stream" > gpurun_out/r2t_tests.txt 2>&1; echo "tests rc=$?" > gpurun_out/r2t_summary.txt
run() { tag=$1; shift; env "$@" MS_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 5 --cpu-sample 0 $EXTRA > gpurun_out/r2t_$tag.json 2> gpurun_out/r2t_$tag.err; echo "$tag rc=$?" >> gpurun_out/r2t_summary.txt; }
run base FOO=1
run thr2 MS_PLAN_THREADS=2
run ramp_d MS_RAMP=64,192,384/256,128,128
run ramp_e MS_RAMP=96,288/256,128,128
tail -3 gpurun_out/r2t_tests.txt; cat gpurun_out/r2t_summary.txt
