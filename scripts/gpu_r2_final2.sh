#!/bin/bash
# 1 GPU, after the tensor-path folds became the default: the default bench line, the ncu launch list of the bench step and a
# full capture of the two round kernels that carry the proof (each after its plain run exited 0); reports exported to CSV
# on the box and removed (gpurun_out/ is limited to 64 MiB).
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err ) 2>&1 | grep real; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2g_bench.json") if l.startswith("{")][-1])
print({k:d[k] for k in ["value","ms_per_step","proof_equals_cpu_oracle_golden","verified","gpu_launches","clocks"]})
print("e2e", d["e2e"]["ms_per_step"]); r=d["roofline"]; print("roofline", {k:r[k] for k in ["achieved","frac","launch_ms","fold_pipe"]}, r["tensor_pipe"]["achieved_imma_per_s"], r["int_pipe"]["frac"])
PY
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-microbench --no-ntt"
$CMD > gpurun_out/r2g_plain.json 2> gpurun_out/r2g_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches.csv $CMD > gpurun_out/r2g_ncu_list.log 2>&1
echo "launch list exit $?"
CMD1="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-microbench --no-ntt"
$CMD1 > gpurun_out/r2g_plain1.json 2> gpurun_out/r2g_plain1.err &&
ncu --set full --clock-control none --import-source on -k regex:round_kernel -s 12 -c 2 -f -o gpurun_out/r2g_round_kernels $CMD1 > gpurun_out/r2g_ncu_full_round.log 2>&1
echo "round kernels capture exit $?"
ncu -i gpurun_out/r2g_round_kernels.ncu-rep --page raw --csv > gpurun_out/r2g_round_kernels_raw.csv 2> /dev/null
ncu -i gpurun_out/r2g_round_kernels.ncu-rep --page source --csv 2> /dev/null | gzip -9 > gpurun_out/r2g_round_kernels_source.csv.gz
rm -f gpurun_out/r2g_round_kernels.ncu-rep
ls -la gpurun_out | grep r2g
