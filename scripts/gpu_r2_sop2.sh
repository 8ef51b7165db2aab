#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sop.py tests/test_gpu_sop_variants.py -m gpu -x -q > gpurun_out/r2_sop2_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2_sop2_pytest.log
for V in ZK_X=1 ZK_B200_SOP_TOOM=0; do
  env $V timeout 120 python scripts/bench_sop.py 24 3 sop > gpurun_out/r2_sop2_bench.jsonl 2>&1
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_sop2_bench.jsonl").readline()); print("[$V]", "prove_ms", round(d["prove_ms"],3), "kernel_ms", round(d["kernel_ms"],3), "first", [round(x,3) for x in d["first_round_ms"]], "gbs", round(d["first_fused_step_gbs"]), "ok", d["verified_against_evaluate"])
PY
done
