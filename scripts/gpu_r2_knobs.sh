#!/bin/bash
# 1 GPU: the parity files once per documented run-time knob, so that no A/B switch ships unchecked
mkdir -p gpurun_out
for V in ZK_B200_SMALL_Q=0 ZK_B200_SMALL_Q=1000000 ZK_B200_H2D_OVERLAP=0 ZK_B200_SCHED=static ZK_B200_FOLD_PIPE=int ZK_B200_FOLD_PIPE=f64 ZK_B200_H2D_STREAMS=1; do
  env $V timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_kats.py tests/test_gpu_sop.py tests/test_gpu_fullsize.py -m gpu -x -q -k "not 2^27 and not 2^28 and not 2^29 and not 2^30" > gpurun_out/r2_knob.log 2>&1; rc=$?
  echo "[$V] exit $rc: $(tail -1 gpurun_out/r2_knob.log)"
  if [ $rc -ne 0 ]; then tail -30 gpurun_out/r2_knob.log; fi
done
