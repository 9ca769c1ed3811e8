#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r3e_summary.txt
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 6 --cpu-sample 0 $EXTRA > gpurun_out/r3e_$tag.json 2> gpurun_out/r3e_$tag.err; echo "$tag rc=$?" >> gpurun_out/r3e_summary.txt; }
EXTRA="--chunk 256" run c256 FOO=1
EXTRA="--chunk 320" run c320 FOO=1
EXTRA="--chunk 384" run c384b FOO=1
cat gpurun_out/r3e_summary.txt
