#!/bin/bash
# round 2, first 1-GPU call: the whole GPU suite with the new full-size NTT cases, then the bench as it stands
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/r2_first_gpu.txt
free -g | head -2 >> gpurun_out/r2_first_gpu.txt; nproc >> gpurun_out/r2_first_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=25 > gpurun_out/r2_first_pytest.log 2>&1; echo "pytest exit $?"
tail -45 gpurun_out/r2_first_pytest.log
timeout 600 python bench.py > gpurun_out/r2_first_bench.json 2> gpurun_out/r2_first_bench.err; echo "bench exit $?"
tail -c 1500 gpurun_out/r2_first_bench.json
timeout 300 python scripts/bench_ntt.py 16 28 > gpurun_out/r2_first_ntt.jsonl 2>&1; echo "ntt exit $?"
cat gpurun_out/r2_first_ntt.jsonl
