#!/bin/bash
# ncu full capture of the first fused launch for both fold pipes
mkdir -p gpurun_out
for pipe in int f64; do
  CMD1="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-microbench"
  ZK_B200_FOLD_PIPE=$pipe ncu --set full --clock-control none --import-source on -k regex:round_kernel -s 1 -c 1 -f -o gpurun_out/prof_ab_${pipe} $CMD1 > gpurun_out/ncu_ab_${pipe}.log 2>&1
  echo "capture $pipe exit $?"
done
ls -la gpurun_out/*.ncu-rep | tail -3
