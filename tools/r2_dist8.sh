#!/bin/bash
set -x
mkdir -p gpurun_out
N=8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  tools/dist_check.py > gpurun_out/r2_dist_check_${N}gpu.txt 2>&1; grep "exchange=\|partitioned build\|ALL OK\|MISMATCH\|Error" gpurun_out/r2_dist_check_${N}gpu.txt | cut -c1-330
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_products_n${N}.json 2> gpurun_out/r2_bench_products_n${N}.err
grep -v Warn gpurun_out/r2_bench_products_n${N}.err | tail -5 | cut -c1-300; head -c 500 gpurun_out/r2_bench_products_n${N}.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 \
  bench.py --gpus $N --workload papers --steps 5 --warmup 3 > gpurun_out/r2_bench_papers_n${N}.json 2> gpurun_out/r2_bench_papers_n${N}.err
grep -v Warn gpurun_out/r2_bench_papers_n${N}.err | tail -8 | cut -c1-300; head -c 500 gpurun_out/r2_bench_papers_n${N}.json
