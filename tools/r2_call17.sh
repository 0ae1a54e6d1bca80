#!/bin/bash
# round 2, call 17: fast-path drain of the TMA-fed dW product: parity, then its time on the products model's shapes
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "gemm or mm or golden or layer_vs_oracle" > gpurun_out/r2_pytest_gemm.txt 2>&1; tail -2 gpurun_out/r2_pytest_gemm.txt
timeout 300 python tools/check_gemm_tc.py > gpurun_out/r2_check_gemm_tc.txt 2>&1; tail -2 gpurun_out/r2_check_gemm_tc.txt
timeout 300 python tools/gemm_tn_probe.py > gpurun_out/r2_gemm_tn_probe_fast_drain_v2.txt 2>&1; grep "R=" gpurun_out/r2_gemm_tn_probe_fast_drain_v2.txt
