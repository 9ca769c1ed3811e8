#!/bin/bash
# round 2, call L: ncu launch list of exactly the timed step (profiler range), --set full of the kernels that carry the step,
# small-config latencies with the plan cache
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plan_cache or c1_c2 or edge_cases" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2l_summary.txt
tail -3 gpurun_out/r2l_pytest.log
for c in C1b C1 C2 C3; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --e2e-steps 20 > gpurun_out/r2l_bench_$c.json 2> gpurun_out/r2l_bench_$c.err; echo "bench $c rc=$?" >> gpurun_out/r2l_summary.txt
done
CMD="python bench.py --renders 512 --steps 1 --warmup 3 --e2e-steps 0 --cpu-sample 0"
$CMD > gpurun_out/r2l_plain.json 2> gpurun_out/r2l_plain.err
rc=$?; echo "plain rc=$rc" >> gpurun_out/r2l_summary.txt
if [ $rc -eq 0 ]; then
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2l_launches_512renders.csv $CMD > gpurun_out/r2l_ncu_list.log 2>&1
  echo "ncu list rc=$?" >> gpurun_out/r2l_summary.txt
  i=0
  for pat in "FirP2K" "ColsWarpK" "RowsK" "SynthNormalK"; do
    i=$((i+1))
    cnt=1; if [ "$pat" = "ColsWarpK" ]; then cnt=4; fi; if [ "$pat" = "RowsK" ]; then cnt=20; fi
    timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$pat" -c $cnt -f -o /tmp/r2l_prof_$i $CMD > gpurun_out/r2l_ncu_$i.log 2>&1
    echo "ncu $i ($pat) rc=$?" >> gpurun_out/r2l_summary.txt
    ncu -i /tmp/r2l_prof_$i.ncu-rep --page details --csv > gpurun_out/r2l_details_$i.csv 2>/dev/null
    ncu -i /tmp/r2l_prof_$i.ncu-rep --page source --csv > gpurun_out/r2l_source_$i.csv 2>/dev/null
  done
fi
du -sh gpurun_out
cat gpurun_out/r2l_summary.txt
