#!/bin/bash
# round 2, second GPU call: the streaming SpMM -- parity tests, then the tuning sweep on the products-shaped R-MAT graph
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream or widths_and_long or bf16_panel_equals" > gpurun_out/r2_pytest_stream.log 2>&1
tail -5 gpurun_out/r2_pytest_stream.log
timeout 500 python tools/rmat_probe.py --widths 256,100,48 --bf16 --fwd-only --reps 3 --check \
  --sweep 0:0:0:0,2:48:2:0,2:48:0:0,2:48:1:0,2:24:2:0,2:96:2:0,2:48:2:16,2:48:2:32 > gpurun_out/r2_rmat_probe_stream.txt 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed
timeout 400 ncu --metrics $M -k regex:spmm_ --clock-control none --csv --log-file gpurun_out/r2_ncu_rmat_stream.csv \
  python tools/rmat_probe.py --widths 256,100 --fwd-only --reps 1 --sweep 2:48:2:0,2:48:0:0 > gpurun_out/r2_ncu_rmat_stream.log 2>&1
