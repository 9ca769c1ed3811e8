#!/bin/bash
# N=$1 ranks: the default bench (kernels + e2e), then kernels-only variants (class chains on one stream; NCCL gather)
set -x
N=${1:-8}
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 400 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 25 --warmup 3 $EXTRA > gpurun_out/r2u_n${N}_$tag.json 2> gpurun_out/r2u_n${N}_$tag.err; echo "$tag rc=$?" >> gpurun_out/r2u_summary_n$N.txt; }
EXTRA="--cpu-sample 0" run default MS_TRACE=1
EXTRA="--e2e-steps 0 --cpu-sample 0" run streams0 MS_SPEC_STREAMS=0
EXTRA="--e2e-steps 0 --cpu-sample 0 --collective gather" run nccl_gather FOO=1
cat gpurun_out/r2u_summary_n$N.txt
