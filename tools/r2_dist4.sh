#!/bin/bash
# round 2: 4 GPUs, products strong scaling with the final exchange code
set -x
mkdir -p gpurun_out
N=4
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_products_n${N}.json 2> gpurun_out/r2_bench_products_n${N}.err
grep -v Warn gpurun_out/r2_bench_products_n${N}.err | tail -5 | cut -c1-300; head -c 300 gpurun_out/r2_bench_products_n${N}.json
