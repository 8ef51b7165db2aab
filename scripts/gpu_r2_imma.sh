#!/bin/bash
# 1 GPU: the tensor-path folds (default for the degree-3, three-factor fused step) against the other fold pipes: the parity
# files and the 2^26 oracle digest per setting (ZK_B200_SMALL_Q=0 sends every round of 32 items or more through the
# streaming kernel, i.e. through the tensor-path folds at every size), then what the driver runs at round end.
mkdir -p gpurun_out
for V in ZK_X=default ZK_B200_SMALL_Q=0 ZK_B200_FOLD_PIPE=imma ZK_B200_FOLD_PIPE=f64 ZK_B200_FOLD_PIPE=int; do
  env $V timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_kats.py tests/test_gpu_sop.py -m gpu -x -q > gpurun_out/r2_imma_knob.log 2>&1; rc=$?
  echo "[$V] pytest exit $rc: $(tail -1 gpurun_out/r2_imma_knob.log)"
  if [ $rc -ne 0 ]; then tail -30 gpurun_out/r2_imma_knob.log; fi
  env $V timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-ntt --no-microbench > gpurun_out/r2_imma_bench.json 2> gpurun_out/r2_imma_bench.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_imma_bench.json") if l.startswith("{")][-1])
    print("[$V] bench ms", round(d["ms_per_step"],3), "golden", d["proof_equals_cpu_oracle_golden"], "verified", d["verified"], "rounds", [round(x,3) for x in d["round_kernel_ms"][:4]])
except Exception as e:
    print("[$V] bench ERR", e)
PY
done
bash scripts/gpu_r2_final1.sh
