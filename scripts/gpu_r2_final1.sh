#!/bin/bash
# 1 GPU, what the driver runs at round end: GPU suite, smoke, the default bench line, the reference arm
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?"
tail -10 gpurun_out/r2f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2f_smoke.log
( time timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err ) 2>&1 | grep real; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2f_bench.json") if l.startswith("{")][-1])
print({k:d[k] for k in ["value","ms_per_step","proof_equals_cpu_oracle_golden","verified","gpu_launches","clocks"]})
print("e2e", d["e2e"]); print("roofline", {k:d["roofline"][k] for k in ["achieved","frac","launch_ms","traffic","traffic_source"]})
n=d["ntt"]; print("ntt", n.get("error") or (n["value"], n["headline_ms"], n["roofline"]["frac"], n["e2e"]["ms"]))
PY
tail -2 gpurun_out/r2f_bench.err
( time timeout 900 python bench.py --impl reference > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err ) 2>&1 | grep real; echo "ref exit $?"; cut -c1-300 gpurun_out/r2f_ref.json
