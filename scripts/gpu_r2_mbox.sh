#!/bin/bash
# N GPUs: the in-kernel mailbox all-reduce — parity under torchrun, then bench with the mailboxes and with the NCCL fallback
N=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_parity.py > gpurun_out/r2_mbox_parity_$N.log 2>&1; echo "dist_parity exit $?"
grep -E "dist_parity|MISMATCH|mailbox|Error|error" gpurun_out/r2_mbox_parity_$N.log | tail -8
for MB in 1 0; do
  ZK_B200_MAILBOX=$MB timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$MB bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-c4 --no-microbench > gpurun_out/r2_mbox_bench_${N}_mb$MB.json 2> gpurun_out/r2_mbox_bench_${N}_mb$MB.err; echo "bench mailbox=$MB exit $?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_mbox_bench_${N}_mb$MB.json") if l.startswith("{")][-1])
    print("mailbox=$MB", "uses_mailbox", d.get("uses_mailbox"), "ms_per_step", round(d["ms_per_step"],4), "golden", d["proof_equals_cpu_oracle_golden"], "verified", d["verified"], "rounds", d["round_kernel_ms"][8:20])
except Exception as e:
    print("ERR", e)
PY
  tail -3 gpurun_out/r2_mbox_bench_${N}_mb$MB.err
done
