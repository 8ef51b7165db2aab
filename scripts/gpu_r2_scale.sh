#!/bin/bash
# N GPUs, exactly what the driver's scaling step runs: the reference arm and our arm under torchrun
N=${1:-4}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2s_ref_$N.json 2> gpurun_out/r2s_ref_$N.err; echo "reference arm exit $?"
cut -c1-200 gpurun_out/r2s_ref_$N.json
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2s_bench_$N.json 2> gpurun_out/r2s_bench_$N.err ) 2>&1 | grep real; echo "bench exit $?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2s_bench_$N.json") if l.startswith("{")][-1])
print({k:d[k] for k in ["n_gpus","value","ms_per_step","proof_equals_cpu_oracle_golden","verified","gpu_launches","uses_mailbox"]})
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["h2d_gbs"]); c=d["config4"]; print("config4", c["prove_ms"], c["value"], c["proof_equals_cpu_oracle_golden"], c["verified"])
PY
tail -3 gpurun_out/r2s_bench_$N.err
