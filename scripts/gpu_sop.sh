#!/bin/bash
# One short gpurun call: the whole GPU suite (not stopping at the first failure), smoke, the sum-of-products timing.
mkdir -p gpurun_out
timeout 60 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_sop.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_sop.log
tail -40 gpurun_out/pytest_gpu_sop.log
timeout 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 20 python scripts/bench_sop.py 24 3 > gpurun_out/bench_sop.jsonl 2> gpurun_out/bench_sop.err; echo "bench_sop exit $?"
cat gpurun_out/bench_sop.jsonl; tail -3 gpurun_out/bench_sop.err
