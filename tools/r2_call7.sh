#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu --maxfail=5 -q > gpurun_out/r2_pytest_gpu_all3.log 2>&1; tail -8 gpurun_out/r2_pytest_gpu_all3.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_products_v2.json 2> gpurun_out/r2_bench_products_v2.err; tail -c 400 gpurun_out/r2_bench_products_v2.err
timeout 300 ncu --set full --import-source on --clock-control none -k regex:spmm_stream_kernel -c 1 -o gpurun_out/r2_spmm_stream_products_f256 \
  python tools/rmat_probe.py --widths 256 --fwd-only --reps 1 --cpu-edges > gpurun_out/r2_ncu_full_stream.log 2>&1
ls -la gpurun_out/*.ncu-rep
