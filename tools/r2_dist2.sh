#!/bin/bash
# 2-GPU call: parity of every exchange / association / split with the single-GPU layer, then the products bench at N=2
set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  tools/dist_check.py > gpurun_out/r2_dist_check_${N}gpu.txt 2>&1; tail -25 gpurun_out/r2_dist_check_${N}gpu.txt | cut -c1-260
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_products_n${N}.json 2> gpurun_out/r2_bench_products_n${N}.err
tail -c 1500 gpurun_out/r2_bench_products_n${N}.err; head -c 3000 gpurun_out/r2_bench_products_n${N}.json
if [ "$N" = 2 ]; then
  timeout 200 python tools/debug_ref_models.py > gpurun_out/r2_debug_ref_models.txt 2>&1; tail -12 gpurun_out/r2_debug_ref_models.txt | cut -c1-200
fi
