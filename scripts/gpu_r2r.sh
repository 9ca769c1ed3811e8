#!/bin/bash
# N=1: class chains on side streams A/B, end-to-end with the new slicing / planner defaults
set -x
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.txt 2>&1; echo "smoke rc=$?" > gpurun_out/r2r_summary.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "spectral or canonical or c5 or planner or batch or stream or cache or graph" > gpurun_out/r2r_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/r2r_summary.txt
run() { tag=$1; shift; env "$@" MS_TRACE=1 timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 5 --cpu-sample 0 $EXTRA > gpurun_out/r2r_$tag.json 2> gpurun_out/r2r_$tag.err; echo "$tag rc=$?" >> gpurun_out/r2r_summary.txt; }
run streams1 MS_SPEC_STREAMS=1
run streams0 MS_SPEC_STREAMS=0
EXTRA="--chunk 384" run chunk384 MS_SPEC_STREAMS=1
tail -3 gpurun_out/r2r_tests.txt; cat gpurun_out/r2r_summary.txt
