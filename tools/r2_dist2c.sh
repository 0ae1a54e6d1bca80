#!/bin/bash
set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  tools/dist_check.py > gpurun_out/r2_dist_check_${N}gpu.txt 2>&1; grep "exchange=\|partitioned build\|ALL OK\|MISMATCH\|Error" gpurun_out/r2_dist_check_${N}gpu.txt | cut -c1-330
for ch in default 16 32; do
  if [ $ch = default ]; then unset NCCL_MIN_P2P_NCHANNELS; else export NCCL_MIN_P2P_NCHANNELS=$ch; export NCCL_MAX_P2P_NCHANNELS=64; fi
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    tools/halo_probe.py 256 8 > gpurun_out/r2_halo_probe_n${N}_ch${ch}.txt 2>&1; grep "world=\|NCCL_M\|Error" gpurun_out/r2_halo_probe_n${N}_ch${ch}.txt | cut -c1-250
done
unset NCCL_MIN_P2P_NCHANNELS NCCL_MAX_P2P_NCHANNELS
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_products_n${N}.json 2> gpurun_out/r2_bench_products_n${N}.err
grep -v Warn gpurun_out/r2_bench_products_n${N}.err | tail -5 | cut -c1-300; head -c 600 gpurun_out/r2_bench_products_n${N}.json
