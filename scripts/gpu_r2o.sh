#!/bin/bash
# gather variants at N=$1: correctness of the three collectives, then the kernels-only bench under each
set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/check_gather_modes.py > gpurun_out/r2o_check_n$N.txt 2>&1; echo "check rc=$?" >> gpurun_out/r2o_summary_n$N.txt
run() { tag=$1; shift; timeout 400 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 25 --warmup 3 --e2e-steps 0 --cpu-sample 0 $EXTRA > gpurun_out/r2o_n${N}_$tag.json 2> gpurun_out/r2o_n${N}_$tag.err; echo "$tag rc=$?" >> gpurun_out/r2o_summary_n$N.txt; }
EXTRA="--collective peer_copy" run peercopy FOO=1
EXTRA="" run gather_memcpy NCCL_P2P_USE_CUDA_MEMCPY=1
EXTRA="" run gather_ch2 NCCL_MAX_P2P_NCHANNELS=2
EXTRA="" run gather_cta4 NCCL_MAX_CTAS=4
EXTRA="" run gather FOO=1
tail -5 gpurun_out/r2o_check_n$N.txt
cat gpurun_out/r2o_summary_n$N.txt
