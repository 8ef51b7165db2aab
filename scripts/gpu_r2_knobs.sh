#!/bin/bash
# 1 GPU: the parity files and the 2^26 oracle digest once per documented run-time knob, so that no A/B
# switch ships unchecked
mkdir -p gpurun_out
for V in ZK_B200_SMALL_Q=0 ZK_B200_SMALL_Q=1000000 ZK_B200_H2D_OVERLAP=0 ZK_B200_SCHED=static ZK_B200_FOLD_PIPE=int ZK_B200_FOLD_PIPE=f64 ZK_B200_H2D_STREAMS=1; do
  env $V timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_kats.py tests/test_gpu_sop.py -m gpu -x -q > gpurun_out/r2_knob.log 2>&1; rc=$?
  echo "[$V] pytest exit $rc: $(tail -1 gpurun_out/r2_knob.log)"
  if [ $rc -ne 0 ]; then tail -30 gpurun_out/r2_knob.log; fi
  env $V timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-ntt --no-microbench > gpurun_out/r2_knob_bench.json 2> gpurun_out/r2_knob_bench.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_knob_bench.json") if l.startswith("{")][-1])
    print("[$V] bench ms", round(d["ms_per_step"],3), "golden", d["proof_equals_cpu_oracle_golden"], "verified", d["verified"], "e2e ms", round(d["e2e"]["ms_per_step"],1))
except Exception as e:
    print("[$V] bench ERR", e)
PY
done
