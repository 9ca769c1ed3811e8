#!/bin/bash
# compute-sanitizer, ONE tool per call: TOOL=racecheck|memcheck
set -x
mkdir -p gpurun_out
TOOL=${TOOL:-racecheck}
timeout 300 python scripts/sanitize_smoke.py > gpurun_out/r2j_plain_$TOOL.log 2>&1; rc=$?; echo "plain rc=$rc" > gpurun_out/r2j_summary_$TOOL.txt
if [ $rc -eq 0 ]; then
  timeout 1500 compute-sanitizer --tool $TOOL --print-limit 50 python scripts/sanitize_smoke.py > gpurun_out/r2j_sanitizer_$TOOL.log 2>&1; echo "$TOOL rc=$?" >> gpurun_out/r2j_summary_$TOOL.txt
  tail -15 gpurun_out/r2j_sanitizer_$TOOL.log
fi
cat gpurun_out/r2j_summary_$TOOL.txt
