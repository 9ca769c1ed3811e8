#!/bin/bash
# multi-GPU run: N=$1
set -x
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 25 --warmup 3 --e2e-steps 3 --cpu-sample 0 > gpurun_out/r2m_bench_n$N.json 2> gpurun_out/r2m_bench_n$N.err; echo "bench n$N rc=$?" > gpurun_out/r2m_summary_n$N.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 25 --warmup 3 --e2e-steps 0 --cpu-sample 0 --no-pipeline > gpurun_out/r2m_bench_n${N}_nopipeline.json 2> gpurun_out/r2m_bench_n${N}_nopipeline.err; echo "bench n$N nopipeline rc=$?" >> gpurun_out/r2m_summary_n$N.txt
cat gpurun_out/r2m_summary_n$N.txt
