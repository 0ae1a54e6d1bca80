#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/train_cora.py > gpurun_out/r2_train_cora.txt 2>&1; tail -4 gpurun_out/r2_train_cora.txt | cut -c1-200
timeout 1500 python -m pytest tests -m gpu --maxfail=8 -q > gpurun_out/r2_pytest_gpu_all6.log 2>&1; tail -6 gpurun_out/r2_pytest_gpu_all6.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_products_v5.json 2> gpurun_out/r2_bench_products_v5.err; tail -c 300 gpurun_out/r2_bench_products_v5.err
