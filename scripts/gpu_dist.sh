#!/bin/bash
# gpurun --gpus N: multi-GPU parity + a short sharded bench.  Usage: bash scripts/gpu_dist.sh N [bench-args]
N=${1:-2}; shift
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_parity.py > gpurun_out/dist_parity_$N.log 2>&1; echo "dist_parity exit $?"
tail -5 gpurun_out/dist_parity_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 "$@" > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err; echo "bench exit $?"
tail -c 2500 gpurun_out/bench_$N.json; tail -5 gpurun_out/bench_$N.err
