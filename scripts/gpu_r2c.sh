#!/bin/bash
# round 2, call C: bench of phase-2 v3, then ncu --set full of the kernels that carry the step
set -x
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; echo "smoke rc=$?" > gpurun_out/r2c_summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --cpu-sample 0 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?" >> gpurun_out/r2c_summary.txt
CMD="python bench.py --renders 256 --steps 1 --warmup 1 --e2e-steps 0 --cpu-sample 0"
$CMD > gpurun_out/r2c_plain.json 2> gpurun_out/r2c_plain.err
rc=$?; echo "plain rc=$rc" >> gpurun_out/r2c_summary.txt
if [ $rc -eq 0 ]; then
  i=0
  for pat in "FirP2K" "FirP1K" "PostMaxK" "ColsK<4, 0, 1, 0, 256>" "ColsK<2, 0, 1, 0, 256>" "RowsK<0, 0, 4, 0>" "SynthNormalK" "OlaK"; do
    i=$((i+1))
    timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$pat" -s 3 -c 1 -f -o gpurun_out/r2c_prof_$i $CMD > gpurun_out/r2c_ncu_$i.log 2>&1
    echo "ncu $i ($pat) rc=$?" >> gpurun_out/r2c_summary.txt
  done
fi
cat gpurun_out/r2c_summary.txt
