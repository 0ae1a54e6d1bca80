#!/bin/bash
# round 2, fifth GPU call: smoke, the whole GPU suite, the first products-shaped bench line
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -3 gpurun_out/r2_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_products_v1.json 2> gpurun_out/r2_bench_products_v1.err; tail -c 600 gpurun_out/r2_bench_products_v1.err
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_all.log 2>&1; tail -5 gpurun_out/r2_pytest_gpu_all.log
