#!/bin/bash
# parity, then A/B over env-selected kernel variants.  VARIANTS="name:ENV=..,ENV=.. ..."  SHAPES="MxD ..."
mkdir -p gpurun_out
if [ -z "$NO_PYTEST" ]; then timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_kats.py -m gpu -x -q > gpurun_out/pytest_ab.log 2>&1; echo "pytest(default) exit $?"
tail -3 gpurun_out/pytest_ab.log; fi
for shape in ${SHAPES:-3x3}; do set -- ${shape%x*} ${shape#*x}
  for var in $VARIANTS; do name=${var%%:*}; envs=$(echo ${var#*:} | tr ',' ' ')
    env $envs timeout 300 python bench.py --steps 5 --warmup 3 --m $1 --degree $2 --no-cpu --no-e2e --no-microbench > gpurun_out/ab_${name}_$1$2.json 2> gpurun_out/ab_${name}_$1$2.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_${name}_$1$2.json").read().strip().splitlines()[-1])
    print("${name} m=$1 D=$2", "ms", round(d["ms_per_step"], 3), "verified", d.get("verified"), "keccak", d.get("proof_keccak", "")[:12], "rounds", d.get("round_kernel_ms", [])[:4], "clk", d.get("clocks", {}).get("sm_mhz"), d.get("clocks", {}).get("reasons"))
except Exception as e:
    print("${name} $1 $2 ERR", e); print(open("gpurun_out/ab_${name}_$1$2.err").read()[-1500:])
PY
  done
done
