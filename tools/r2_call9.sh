#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu --maxfail=8 -q > gpurun_out/r2_pytest_gpu_all5.log 2>&1; tail -8 gpurun_out/r2_pytest_gpu_all5.log | cut -c1-200
timeout 400 python tools/rmat_probe.py --widths 256,100 --relabel none,rows,degree --fwd-only --reps 3 --cpu-edges --sweep 1:48:0:0:0,1:48:0:0:1 > gpurun_out/r2_rmat_probe_roworder.txt 2>&1; grep "relabel=\| f=\|config" gpurun_out/r2_rmat_probe_roworder.txt | cut -c1-160
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_products_v4.json 2> gpurun_out/r2_bench_products_v4.err; tail -c 300 gpurun_out/r2_bench_products_v4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_products.csv \
  python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
