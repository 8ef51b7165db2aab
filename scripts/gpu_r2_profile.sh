#!/bin/bash
# 1 GPU: ncu launch list of the bench step + full captures of the kernels that carry the time (each after its plain run
# exited 0).  The .ncu-rep files are exported to CSV on the box and removed (gpurun_out/ is limited to 64 MiB).
mkdir -p gpurun_out
exp() {  # name: raw page (+ source page, gzipped) of gpurun_out/NAME.ncu-rep, then drop the report
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2> /dev/null
  if [ "$2" = src ]; then ncu -i gpurun_out/$1.ncu-rep --page source --csv 2> /dev/null | gzip -9 > gpurun_out/$1_source.csv.gz; fi
  rm -f gpurun_out/$1.ncu-rep
}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-microbench --no-ntt"
$CMD > gpurun_out/r2p_plain.json 2> gpurun_out/r2p_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_launches.csv $CMD > gpurun_out/r2p_ncu_list.log 2>&1
echo "launch list exit $?"
CMD1="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-microbench --no-ntt"
$CMD1 > gpurun_out/r2p_plain1.json 2> gpurun_out/r2p_plain1.err &&
ncu --set full --clock-control none --import-source on -k regex:round_kernel -s 12 -c 2 -f -o gpurun_out/r2p_round_kernels $CMD1 > gpurun_out/r2p_ncu_full_round.log 2>&1
echo "round kernels capture exit $?"; exp r2p_round_kernels src
ncu --set full --clock-control none -k regex:round_small_kernel -s 14 -c 2 -f -o gpurun_out/r2p_small_kernel $CMD1 > gpurun_out/r2p_ncu_full_small.log 2>&1
echo "small kernel capture exit $?"; exp r2p_small_kernel
SOP="python scripts/bench_sop.py 24 1 sop"
$SOP > gpurun_out/r2p_sop_plain.jsonl 2>&1 &&
ncu --set full --clock-control none -k regex:sop_round_kernel -s 10 -c 2 -f -o gpurun_out/r2p_sop_kernels $SOP > gpurun_out/r2p_ncu_full_sop.log 2>&1
echo "sop capture exit $?"; exp r2p_sop_kernels
NTT="python scripts/bench_ntt.py 24 24"
$NTT > gpurun_out/r2p_ntt_plain.jsonl 2>&1 &&
ncu --set full --clock-control none -k regex:ntt_pass_kernel -s 3 -c 3 -f -o gpurun_out/r2p_ntt_kernels $NTT > gpurun_out/r2p_ncu_full_ntt.log 2>&1
echo "ntt capture exit $?"; exp r2p_ntt_kernels
ls -la gpurun_out | tail -20; du -sh gpurun_out
