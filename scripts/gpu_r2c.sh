#!/bin/bash
# round 2, call C: bench of phase-2 v3, then ncu --set full of three kernels; reports are reduced to CSV pages on the box
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --cpu-sample 0 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?" > gpurun_out/r2c_summary.txt
CMD="python bench.py --renders 256 --steps 1 --warmup 1 --e2e-steps 0 --cpu-sample 0"
$CMD > gpurun_out/r2c_plain.json 2> gpurun_out/r2c_plain.err
rc=$?; echo "plain rc=$rc" >> gpurun_out/r2c_summary.txt
if [ $rc -eq 0 ]; then
  i=0
  for pat in "FirP2K" "PostMaxK" "ColsK<4, 0, 1, 0, 256>" "FirP1K" "SynthNormalK"; do
    i=$((i+1))
    timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$pat" -s 3 -c 1 -f -o /tmp/r2c_prof_$i $CMD > gpurun_out/r2c_ncu_$i.log 2>&1
    echo "ncu $i ($pat) rc=$?" >> gpurun_out/r2c_summary.txt
    ncu -i /tmp/r2c_prof_$i.ncu-rep --page details --csv > gpurun_out/r2c_details_$i.csv 2>/dev/null
    ncu -i /tmp/r2c_prof_$i.ncu-rep --page source --csv > gpurun_out/r2c_source_$i.csv 2>/dev/null
    ncu -i /tmp/r2c_prof_$i.ncu-rep --page raw --csv > gpurun_out/r2c_raw_$i.csv 2>/dev/null
  done
fi
du -sh gpurun_out
cat gpurun_out/r2c_summary.txt
