#!/bin/bash
# One gpurun call: GPU tests, smoke, a short bench.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu_info.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
