#!/bin/bash
# 1 GPU: the default bench line (with the NTT record) and the reference arm, as the driver runs them
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r2_bench_1.json 2> gpurun_out/r2_bench_1.err ) 2>&1 | grep real; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_1.json") if l.startswith("{")][-1])
print({k:d[k] for k in ["value","ms_per_step","proof_equals_cpu_oracle_golden","verified","gpu_launches"]}, d["e2e"])
print("roofline", {k:d["roofline"][k] for k in ["achieved","frac","launch_ms","traffic"]})
n=d["ntt"]
if "error" in n: print("NTT ERROR", n)
else:
    print("ntt", n["value"], n["headline_ms"], n["roofline"]["frac"], n["e2e"], n["cpu_baseline"])
    for r in n["sweep"]: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k in ("field","log_n","ntt_ms","intt_ms","butterfly_mul_per_s","int_pipe_frac","fft_equals_cpu_oracle_golden","ntt_first_call_ms_incl_plan_build")})
PY
tail -3 gpurun_out/r2_bench_1.err
( time timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err ) 2>&1 | grep real; echo "ref exit $?"
cut -c1-1500 gpurun_out/r2_bench_ref.json
