#!/bin/bash
# 1 GPU: the sum-of-products kernel variants — parity file per variant, then the GKR 4 x 2^24 timing
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m pytest tests/test_gpu_sop.py -m gpu -q -p no:cacheprovider > gpurun_out/r2_sop_pytest_$name.log 2>&1; echo "pytest($name) exit $? $(tail -1 gpurun_out/r2_sop_pytest_$name.log)"
  env "$@" timeout 120 python scripts/bench_sop.py 24 3 sop > gpurun_out/r2_sop_bench_$name.jsonl 2> gpurun_out/r2_sop_bench_$name.err; echo "bench($name) exit $?"
  python - <<PY
import json
for l in open("gpurun_out/r2_sop_bench_$name.jsonl"):
    d=json.loads(l); print("$name", d["poly"], "prove_ms", round(d["prove_ms"],3), "kernel_ms", round(d["kernel_ms"],3), "first", [round(x,3) for x in d["first_round_ms"]], "ok", d["verified_against_evaluate"])
PY
  tail -2 gpurun_out/r2_sop_bench_$name.err
}
run default ZK_X=1
run notoom ZK_B200_SOP_TOOM=0
run nogroup ZK_B200_SOP_GROUP=0
run static ZK_B200_SOP_SCHED=static
run narrow ZK_B200_SOP_WIDE=0
run narrow_nogroup_static ZK_B200_SOP_WIDE=0 ZK_B200_SOP_GROUP=0 ZK_B200_SOP_SCHED=static
run f64 ZK_B200_SOP_FOLD_PIPE=f64
timeout 120 python scripts/bench_sop.py 24 3 > gpurun_out/r2_sop_bench_with_product.jsonl 2>&1; cat gpurun_out/r2_sop_bench_with_product.jsonl | cut -c1-400
