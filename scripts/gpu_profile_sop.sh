#!/bin/bash
# One gpurun call (1 GPU): the sum-of-products timing, its ncu launch list, one full capture of the fused kernel.
# Usage: bash scripts/gpu_profile_sop.sh <tag> [log_n=24]
TAG=${1:-r02}; N=${2:-24}
mkdir -p gpurun_out
CMD="python scripts/bench_sop.py $N 2 sop"
$CMD > gpurun_out/sop_plain_${TAG}.jsonl 2> gpurun_out/sop_plain_${TAG}.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/sop_launches_${TAG}.csv $CMD > gpurun_out/sop_ncu_list_${TAG}.log 2>&1
echo "launch list exit $?"
# launch 1 of the prover = round 0, launch 2 = the first fused fold+sum step (after the claim's launch): capture both
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sop_round_kernel -s 1 -c 2 -f -o gpurun_out/sop_prof_${TAG} $CMD > gpurun_out/sop_ncu_full_${TAG}.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out | tail -8
