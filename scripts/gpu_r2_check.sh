#!/bin/bash
# 1 GPU: whole GPU suite, default bench line, e2e A/B of the overlapped upload, sum-of-products timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2_check_pytest.log 2>&1; echo "pytest exit $?"
tail -14 gpurun_out/r2_check_pytest.log
for OV in 1 0; do
  ZK_B200_H2D_OVERLAP=$OV timeout 600 python bench.py --no-ntt --no-cpu > gpurun_out/r2_check_bench_ov$OV.json 2> gpurun_out/r2_check_bench_ov$OV.err; echo "bench overlap=$OV exit $?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_check_bench_ov$OV.json") if l.startswith("{")][-1])
print("overlap=$OV", {k:d[k] for k in ["value","ms_per_step","proof_equals_cpu_oracle_golden","verified","gpu_launches"]}, "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], d["e2e"]["h2d_gbs"])
PY
  tail -2 gpurun_out/r2_check_bench_ov$OV.err
done
timeout 120 python scripts/bench_sop.py 24 3 sop > gpurun_out/r2_check_sop.jsonl 2>&1; cut -c1-420 gpurun_out/r2_check_sop.jsonl
