#!/bin/bash
# 1 GPU, final state of round 2: the parity files with every round of 32 items or more on the streaming (tensor-path) kernel,
# what the driver runs at round end (suite, smoke, default bench line, reference arm), then the ncu launch list of the bench
# step and a full capture of the two round kernels that carry the proof (reports exported to CSV on the box and removed).
mkdir -p gpurun_out
ZK_B200_SMALL_Q=0 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_kats.py tests/test_gpu_sop.py -m gpu -x -q > gpurun_out/r2h_smallq0.log 2>&1; echo "[ZK_B200_SMALL_Q=0] pytest exit $?: $(tail -1 gpurun_out/r2h_smallq0.log)"
bash scripts/gpu_r2_final1.sh
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-microbench --no-ntt"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches.csv $CMD > gpurun_out/r2g_ncu_list.log 2>&1
echo "launch list exit $?"
CMD1="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-microbench --no-ntt"
ncu --set full --clock-control none --import-source on -k regex:round_kernel -s 12 -c 2 -f -o gpurun_out/r2g_round_kernels $CMD1 > gpurun_out/r2g_ncu_full_round.log 2>&1
echo "round kernels capture exit $?"
ncu -i gpurun_out/r2g_round_kernels.ncu-rep --page raw --csv > gpurun_out/r2g_round_kernels_raw.csv 2> /dev/null
ncu -i gpurun_out/r2g_round_kernels.ncu-rep --page source --csv 2> /dev/null | gzip -9 > gpurun_out/r2g_round_kernels_source.csv.gz
rm -f gpurun_out/r2g_round_kernels.ncu-rep
