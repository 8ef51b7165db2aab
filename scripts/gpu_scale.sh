#!/bin/bash
# gpurun --gpus N -- bash scripts/gpu_scale.sh N : parity + default weak-scaling bench (+ C4 at N=8)
N=${1:-8}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_parity.py > gpurun_out/dist_parity_$N.log 2>&1; echo "dist_parity exit $?"
grep dist_parity gpurun_out/dist_parity_$N.log | tail -n 1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err; echo "bench exit $?"
if [ "$N" = "8" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --log-n 30 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_c4_$N.json 2> gpurun_out/bench_c4_$N.err; echo "bench c4 exit $?"
fi
python - <<PY
import json, glob
for f in ("gpurun_out/bench_$N.json","gpurun_out/bench_c4_$N.json"):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, {k:d.get(k) for k in ["value","ms_per_step","verified","proof_keccak","step_ms","gpu_launches","n_gpus"]}, d["config"]["log_n"], d.get("e2e"))
    except Exception as e: print(f, "ERR", e)
PY
