#!/bin/bash
# round 2: 8 GPUs, products strong scaling: the default (pipelined exchange in 2 column chunks, high-priority side stream,
# panels produced in the compact buffer) with parity gate, and the unchunked exchange for comparison (quick)
set -x
mkdir -p gpurun_out
N=8
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_products_n${N}.json 2> gpurun_out/r2_bench_products_n${N}.err
grep -v Warn gpurun_out/r2_bench_products_n${N}.err | tail -5 | cut -c1-300; head -c 300 gpurun_out/r2_bench_products_n${N}.json
GCNB_DIST_CHUNKS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 \
  bench.py --gpus $N --steps 8 --warmup 3 --quick > gpurun_out/r2_tune_n${N}_chunks1.json 2> gpurun_out/r2_tune_n${N}_chunks1.err
head -c 300 gpurun_out/r2_tune_n${N}_chunks1.json
