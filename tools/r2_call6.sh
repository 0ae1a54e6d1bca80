#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/layer_ops_probe.py products 256 47 3 > gpurun_out/r2_layer_ops_256_47.txt 2>&1; cat gpurun_out/r2_layer_ops_256_47.txt | grep " ms"
timeout 300 python tools/layer_ops_probe.py products 256 256 3 > gpurun_out/r2_layer_ops_256_256.txt 2>&1; cat gpurun_out/r2_layer_ops_256_256.txt | grep " ms"
timeout 300 python tools/layer_ops_probe.py products 100 256 3 > gpurun_out/r2_layer_ops_100_256.txt 2>&1; cat gpurun_out/r2_layer_ops_100_256.txt | grep " ms"
timeout 1500 python -m pytest tests -m gpu --maxfail=5 -q > gpurun_out/r2_pytest_gpu_all2.log 2>&1; tail -8 gpurun_out/r2_pytest_gpu_all2.log
