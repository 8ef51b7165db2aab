#!/bin/bash
# One gpurun call (1 GPU): plain bench run, then the ncu launch list and one full capture of the dominant kernel.
# Usage: bash scripts/gpu_profile.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-microbench"
$CMD > gpurun_out/plain_${TAG}.json 2> gpurun_out/plain_${TAG}.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "launch list exit $?"
CMD1="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-microbench"
$CMD1 > gpurun_out/plain1_${TAG}.json 2> gpurun_out/plain1_${TAG}.err &&
ncu --set full --clock-control none --import-source on -k regex:round_kernel -s 1 -c 1 -f -o gpurun_out/prof_${TAG} $CMD1 > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out | tail -12
