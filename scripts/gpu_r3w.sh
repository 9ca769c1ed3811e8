#!/bin/bash
# round 2 artefact run at N=1: smoke, both bench arms, the other BASELINE configs, ncu launch list (512-render slab)
set -x
mkdir -p gpurun_out
S=gpurun_out/r3w_summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3w_smoke.log 2>&1; echo "smoke rc=$?" > $S
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r3w_pytest.log 2>&1; echo "pytest rc=$?" >> $S
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3w_bench_ref.json 2> gpurun_out/r3w_bench_ref.err; echo "bench ref rc=$?" >> $S
timeout 900 python bench.py > gpurun_out/r3w_bench_n1.json 2> gpurun_out/r3w_bench_n1.err; echo "bench n1 rc=$?" >> $S
for c in C1b C1 C2 C3 C4; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r3w_bench_$c.json 2> gpurun_out/r3w_bench_$c.err; echo "bench $c rc=$?" >> $S
done
CMD="python bench.py --renders 512 --steps 1 --warmup 3 --e2e-steps 0 --cpu-sample 0"
$CMD > gpurun_out/r3w_plain.json 2> gpurun_out/r3w_plain.err
rc=$?; echo "plain rc=$rc" >> $S
if [ $rc -eq 0 ]; then
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r3w_launches_512renders.csv $CMD > gpurun_out/r3w_ncu_list.log 2>&1
  echo "ncu list rc=$?" >> $S
fi
cat $S
