#!/bin/bash
# One short gpurun call: the sum-of-products GPU tests and timing with integer folds, FP64 folds, deferred reduction.
mkdir -p gpurun_out
for pipe in int f64 wide; do
  unset ZK_B200_SOP_WIDE ZK_B200_SOP_FOLD_PIPE
  if [ $pipe = wide ]; then export ZK_B200_SOP_WIDE=1; else export ZK_B200_SOP_FOLD_PIPE=$pipe; fi
  timeout 60 python -m pytest tests/test_gpu_sop.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_sop_$pipe.log 2>&1; echo "pytest($pipe) exit $?" >> gpurun_out/pytest_sop_$pipe.log
  tail -4 gpurun_out/pytest_sop_$pipe.log
  timeout 60 python scripts/bench_sop.py 24 3 sop > gpurun_out/bench_sop_$pipe.jsonl 2> gpurun_out/bench_sop_$pipe.err; echo "bench_sop($pipe) exit $?"
  cat gpurun_out/bench_sop_$pipe.jsonl; tail -2 gpurun_out/bench_sop_$pipe.err
done
