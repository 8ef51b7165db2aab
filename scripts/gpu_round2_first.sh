#!/bin/bash
# First single-GPU call of the next round: everything written after the round-1 GPU budget was spent.
mkdir -p gpurun_out
timeout 300 python scripts/check_virtual_ntt.py 20 > gpurun_out/virtual_ntt.json 2> gpurun_out/virtual_ntt.err; echo "virtual ntt exit $?"
tail -c 1500 gpurun_out/virtual_ntt.json; tail -3 gpurun_out/virtual_ntt.err
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/pytest_gpu.log
