#!/bin/bash
# round 2, call I: batched tile loads; bench + e2e; ncu of the warp Bluestein columns and the rows kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fft_lengths or spectral_ops or c5_members or golden or c1_c2 or c3" > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2i_summary.txt
tail -3 gpurun_out/r2i_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --e2e-steps 4 --cpu-sample 0 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?" >> gpurun_out/r2i_summary.txt
CMD="python bench.py --renders 512 --steps 1 --warmup 1 --e2e-steps 0 --cpu-sample 0"
$CMD > gpurun_out/r2i_plain.json 2> gpurun_out/r2i_plain.err
rc=$?; echo "plain rc=$rc" >> gpurun_out/r2i_summary.txt
if [ $rc -eq 0 ]; then
  i=0
  for pat in "ColsWarpK<4" "ColsWarpK<2" "RowsK<0, 0, 4, 0>" "ColsK<4, 0, 1, 0, 512>" "OlaK"; do
    i=$((i+1))
    timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$pat" -s 8 -c 1 -f -o /tmp/r2i_prof_$i $CMD > gpurun_out/r2i_ncu_$i.log 2>&1
    echo "ncu $i ($pat) rc=$?" >> gpurun_out/r2i_summary.txt
    ncu -i /tmp/r2i_prof_$i.ncu-rep --page details --csv > gpurun_out/r2i_details_$i.csv 2>/dev/null
    ncu -i /tmp/r2i_prof_$i.ncu-rep --page source --csv > gpurun_out/r2i_source_$i.csv 2>/dev/null
  done
fi
cat gpurun_out/r2i_summary.txt
