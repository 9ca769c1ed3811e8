#!/bin/bash
set -x
N=${1:-8}
mkdir -p gpurun_out
MS_RANK_TIMES=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --e2e-steps 0 --cpu-sample 0 > gpurun_out/r3n_n$N.json 2> gpurun_out/r3n_n$N.err
MS_RANK_TIMES=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --e2e-steps 0 --cpu-sample 0 --no-gather > gpurun_out/r3n_n${N}_nogather.json 2> gpurun_out/r3n_n${N}_nogather.err
grep "^rank" gpurun_out/r3n_n$N.err gpurun_out/r3n_n${N}_nogather.err
