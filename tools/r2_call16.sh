#!/bin/bash
# round 2, call 16: fast-path epilogue of the TMA-fed rows product: parity, then the K / N sweep again
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "gemm or mm or golden or layer_vs_oracle" > gpurun_out/r2_pytest_gemm.txt 2>&1; tail -2 gpurun_out/r2_pytest_gemm.txt
timeout 300 python tools/check_gemm_tc.py > gpurun_out/r2_check_gemm_tc.txt 2>&1; tail -3 gpurun_out/r2_check_gemm_tc.txt
timeout 300 python tools/gemm_k_sweep.py > gpurun_out/r2_gemm_k_sweep_fast_epilogue.txt 2>&1; grep "M=" gpurun_out/r2_gemm_k_sweep_fast_epilogue.txt
