#!/bin/bash
# round 2, call 15: ncu --set full of the two tcgen05 products of the products model's middle layer (2.45 M x 256 x 256)
set -x
mkdir -p gpurun_out
timeout 300 python tools/gemm_one.py rows 2449029 256 256 > gpurun_out/r2_gemm_rows_plain.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tma_kernel -c 1 -f -o gpurun_out/r2_gemm_rows_256 python tools/gemm_one.py rows 2449029 256 256 > gpurun_out/r2_gemm_rows_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tma_kernel -c 1 -f -o gpurun_out/r2_gemm_tn_256 python tools/gemm_one.py tn 256 256 2449029 > gpurun_out/r2_gemm_tn_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/r2_gemm_rows_ncu.log
