#!/bin/bash
# round 2 multi-GPU call: bash scripts/gpu_r2_dist.sh N   (gpurun --gpus N)
# 1. the multi-GPU pytest case (tests/dist_parity.py under torchrun: sharded provers, sum of products, zk_ntt_sharded)
# 2. bench.py at N GPUs (weak scaling + the config-4 record at 2^30)   3. the sharded NTT sweep   4. host-fabric H2D probe
N=${1:-2}
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
nvidia-smi topo -m > gpurun_out/r2_topo_$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_dist_pytest_$N.log 2>&1; echo "pytest multi exit $?"
tail -5 gpurun_out/r2_dist_pytest_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_$N.json 2> gpurun_out/r2_bench_$N.err; echo "bench exit $?"
tail -c 1800 gpurun_out/r2_bench_$N.json; tail -5 gpurun_out/r2_bench_$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 scripts/bench_ntt_sharded.py 20 28 > gpurun_out/r2_ntt_sharded_$N.jsonl 2> gpurun_out/r2_ntt_sharded_$N.err; echo "ntt sharded exit $?"
cat gpurun_out/r2_ntt_sharded_$N.jsonl; tail -3 gpurun_out/r2_ntt_sharded_$N.err
timeout 300 python scripts/h2d_probe.py $N > gpurun_out/r2_h2d_probe_$N.json 2>&1; echo "h2d probe exit $?"
cat gpurun_out/r2_h2d_probe_$N.json
