#!/bin/bash
# ncu --set full of the kernels that carry the step (final code of round 2), 512-render slab; ONE ncu invocation
set -x
mkdir -p gpurun_out
CMD="python bench.py --renders 512 --steps 1 --warmup 3 --e2e-steps 0 --cpu-sample 0"
$CMD > gpurun_out/r3x_plain.json 2> gpurun_out/r3x_plain.err
rc=$?; echo "plain rc=$rc" > gpurun_out/r3x_summary.txt
if [ $rc -eq 0 ]; then
  timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
     -k "regex:FirP1K|FirP2K|FirP3K|PostMaxK|PostWriteK|OlaK|synth_normal_cluster|ColsWarpK<|ColsWarp512K|SpecOpK|SynthDustK|SynthTiltK" -c 30 -f -o /tmp/r3x_prof $CMD > gpurun_out/r3x_ncu.log 2>&1
  echo "ncu rc=$?" >> gpurun_out/r3x_summary.txt
  ncu -i /tmp/r3x_prof.ncu-rep --page details --csv > gpurun_out/r3x_details.csv 2>/dev/null
  ncu -i /tmp/r3x_prof.ncu-rep --page raw --csv > gpurun_out/r3x_raw.csv 2>/dev/null
  ncu -i /tmp/r3x_prof.ncu-rep --page source --csv --kernel-name regex:FirP2K > gpurun_out/r3x_source_p2.csv 2>/dev/null
fi
du -sh gpurun_out; ls -la gpurun_out/r3x_*
cat gpurun_out/r3x_summary.txt
