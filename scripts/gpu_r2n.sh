#!/bin/bash
# collective variants at N=$1 (kernels only)
set -x
N=${1:-8}
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 25 --warmup 3 --e2e-steps 0 --cpu-sample 0 $EXTRA > gpurun_out/r2n_n${N}_$tag.json 2> gpurun_out/r2n_n${N}_$tag.err; echo "$tag rc=$?" >> gpurun_out/r2n_summary_n$N.txt; }
EXTRA="" run gather FOO=1
EXTRA="" run gather_ch32 NCCL_MAX_P2P_NCHANNELS=32 NCCL_MIN_P2P_NCHANNELS=16
EXTRA="--collective all_gather" run allgather FOO=1
cat gpurun_out/r2n_summary_n$N.txt
