#!/bin/bash
# 1 GPU: the latency kernel of the small rounds — parity suite, then the per-round times for several thresholds
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "not ntt and not fullsize_digest" > gpurun_out/r2_small_pytest.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/r2_small_pytest.log
for Q in 0 4096 32768 131072 524288; do
  ZK_B200_SMALL_Q=$Q timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-ntt --no-microbench > gpurun_out/r2_small_q$Q.json 2> gpurun_out/r2_small_q$Q.err; echo "bench q=$Q exit $?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_small_q$Q.json") if l.startswith("{")][-1])
print("Q=$Q", "ms_per_step", round(d["ms_per_step"],4), "golden", d["proof_equals_cpu_oracle_golden"], "rounds", d["round_kernel_ms"][6:])
PY
done
