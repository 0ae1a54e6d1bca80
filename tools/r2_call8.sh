#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 200 python tools/debug_ref_models.py > gpurun_out/r2_debug_ref_models_trunc.txt 2>&1; grep "alone" gpurun_out/r2_debug_ref_models_trunc.txt
GCNB_TC_HI=rna timeout 200 python tools/debug_ref_models.py > gpurun_out/r2_debug_ref_models_rna.txt 2>&1; grep "alone" gpurun_out/r2_debug_ref_models_rna.txt
GCNB_TC_HI=rna timeout 300 python tools/layer_ops_probe.py products 256 256 3 > gpurun_out/r2_layer_ops_256_256_rna.txt 2>&1; grep " ms" gpurun_out/r2_layer_ops_256_256_rna.txt
timeout 1500 python -m pytest tests -m gpu --maxfail=8 -q > gpurun_out/r2_pytest_gpu_all4.log 2>&1; tail -12 gpurun_out/r2_pytest_gpu_all4.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_products_v3.json 2> gpurun_out/r2_bench_products_v3.err; tail -c 300 gpurun_out/r2_bench_products_v3.err
