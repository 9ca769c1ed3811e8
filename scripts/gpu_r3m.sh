#!/bin/bash
# the driver's multi-GPU bench at N=$1 (default flags), plus the reference arm line it prints at N>1
set -x
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r3m_bench_n$N.json 2> gpurun_out/r3m_bench_n$N.err; echo "bench n$N rc=$?" > gpurun_out/r3m_summary_n$N.txt
tail -c 600 gpurun_out/r3m_bench_n$N.json; cat gpurun_out/r3m_summary_n$N.txt
