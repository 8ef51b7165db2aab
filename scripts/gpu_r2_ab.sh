#!/bin/bash
# 1 GPU: A/B of environment-selected kernel variants on the headline bench.  usage: gpu_r2_ab.sh "VAR=val ..." "VAR=val ..." ...
mkdir -p gpurun_out
i=0
for V in "$@"; do
  i=$((i+1))
  env $V timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-ntt --no-microbench > gpurun_out/r2_ab_$i.json 2> gpurun_out/r2_ab_$i.err; rc=$?
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_ab_$i.json") if l.startswith("{")][-1])
    print("[$V] rc=$rc ms_per_step", round(d["ms_per_step"],4), "golden", d["proof_equals_cpu_oracle_golden"], "rounds", d["round_kernel_ms"][:8])
except Exception as e:
    print("[$V] rc=$rc ERR", e)
PY
  tail -2 gpurun_out/r2_ab_$i.err
done
