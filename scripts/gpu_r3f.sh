#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/r3f_summary.txt
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 0 --cpu-sample 0 > gpurun_out/r3f_$tag.json 2> gpurun_out/r3f_$tag.err; echo "$tag rc=$?" >> gpurun_out/r3f_summary.txt; }
run base FOO=1
run w3 MS_LIB_PATH=$PWD/variants/w3/libmicrosound_b200.so
cat gpurun_out/r3f_summary.txt
