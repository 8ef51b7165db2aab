#!/bin/bash
# One gpurun call: parity with the default kernels, then A/B of the staging (ZK_B200_STAGE=reg|tma) and the fold pipe
# (ZK_B200_FOLD_PIPE=int|f64) on three shapes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_kats.py -m gpu -x -q > gpurun_out/pytest_ab.log 2>&1; echo "pytest(default) exit $?"
tail -4 gpurun_out/pytest_ab.log
ZK_B200_FOLD_PIPE=f64 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_ab_f64.log 2>&1; echo "pytest(f64) exit $?"
tail -2 gpurun_out/pytest_ab_f64.log
for shape in ${SHAPES:-"3 3" "2 2" "1 1"}; do set -- $shape
  for stage in reg tma; do for pipe in int f64; do
    ZK_B200_STAGE=$stage ZK_B200_FOLD_PIPE=$pipe timeout 300 python bench.py --steps 5 --warmup 3 --m $1 --degree $2 --no-cpu --no-e2e --no-microbench > gpurun_out/ab_${stage}_${pipe}_$1$2.json 2> gpurun_out/ab_${stage}_${pipe}_$1$2.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_${stage}_${pipe}_$1$2.json").read().strip().splitlines()[-1])
    print("${stage} ${pipe} m=$1 D=$2", "ms", round(d["ms_per_step"], 3), "verified", d.get("verified"), "keccak", d.get("proof_keccak", "")[:12], "rounds", d.get("round_kernel_ms", [])[:4], "clk", d.get("clocks", {}).get("sm_mhz"), d.get("clocks", {}).get("reasons"))
except Exception as e:
    print("${stage} ${pipe} $1 $2 ERR", e); print(open("gpurun_out/ab_${stage}_${pipe}_$1$2.err").read()[-1500:])
PY
  done; done
done
