#!/bin/bash
# round 2, final validation after the epilogue change: full GPU suite, bench line, launch list of the same command
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu_final.log 2>&1; tail -2 gpurun_out/r2_pytest_gpu_final.log
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 300 gpurun_out/r2_bench_final.err; head -c 300 gpurun_out/r2_bench_final.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_products_final.csv \
  python bench.py --steps 2 --warmup 1 --quick > gpurun_out/r2_ncu_launches.log 2>&1; tail -2 gpurun_out/r2_ncu_launches.log | cut -c1-200
