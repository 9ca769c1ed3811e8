#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 120 python scripts/hostplan_timing.py > gpurun_out/r2s_hostplan_timing.txt 2>&1
cat gpurun_out/r2s_hostplan_timing.txt
