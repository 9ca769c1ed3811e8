#!/bin/bash
# N=1 end-to-end: threaded native planner, GIL switch interval, slice ramps (trace on stderr)
set -x
mkdir -p gpurun_out
nproc > gpurun_out/r2q_summary.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "planner or batch or stream" > gpurun_out/r2q_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/r2q_summary.txt
run() { tag=$1; shift; env "$@" MS_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 5 --cpu-sample 0 > gpurun_out/r2q_$tag.json 2> gpurun_out/r2q_$tag.err; echo "$tag rc=$?" >> gpurun_out/r2q_summary.txt; }
run base FOO=1
run ramp_a MS_RAMP=64,128,256/256,128,64
run ramp_b MS_RAMP=32,64,128,256/256,128,128
run ramp_c MS_RAMP=128,384/256,128,128
run thr4 MS_PLAN_THREADS=4
tail -3 gpurun_out/r2q_tests.txt; cat gpurun_out/r2q_summary.txt
