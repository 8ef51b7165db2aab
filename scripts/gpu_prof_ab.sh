#!/bin/bash
# ncu full capture of round 0 and the first fused launch.  Usage: gpu_prof_ab.sh <tag> [env assignments...]
TAG=${1:-x}; shift
mkdir -p gpurun_out
CMD1="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-microbench ${BENCH_ARGS}"
env "$@" ncu --set full --clock-control none --import-source on -k regex:round_kernel -s 26 -c 2 -f -o gpurun_out/prof_${TAG} $CMD1 > gpurun_out/ncu_${TAG}.log 2>&1
echo "capture $TAG exit $?"
ls -la gpurun_out/prof_${TAG}.ncu-rep
