#!/bin/bash
# round 2, call K: the full GPU suite, the driver's N=1 bench (both arms), the ncu launch list with DRAM bytes, small configs
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/r2k_summary.txt
tail -3 gpurun_out/r2k_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2k_summary.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2k_bench_ref.json 2> gpurun_out/r2k_bench_ref.err; echo "bench ref rc=$?" >> gpurun_out/r2k_summary.txt
timeout 900 python bench.py > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err; echo "bench n1 rc=$?" >> gpurun_out/r2k_summary.txt
for c in C1b C1 C2 C3 C4; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2k_bench_$c.json 2> gpurun_out/r2k_bench_$c.err; echo "bench $c rc=$?" >> gpurun_out/r2k_summary.txt
done
CMD="python bench.py --renders 512 --steps 1 --warmup 3 --e2e-steps 0 --cpu-sample 0"
$CMD > gpurun_out/r2k_plain.json 2> gpurun_out/r2k_plain.err
rc=$?; echo "plain rc=$rc" >> gpurun_out/r2k_summary.txt
if [ $rc -eq 0 ]; then
  # one step = 50 launches; 3 warm-up steps + the timed one come first: skip 150, take the timed step
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 150 -c 50 --csv --log-file gpurun_out/r2k_launches_512renders.csv $CMD > gpurun_out/r2k_ncu_list.log 2>&1
  echo "ncu list rc=$?" >> gpurun_out/r2k_summary.txt
fi
cat gpurun_out/r2k_summary.txt
