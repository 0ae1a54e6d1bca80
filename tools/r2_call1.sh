#!/bin/bash
# round 2, first GPU call: everything that had never run on hardware + the products-shaped R-MAT SpMM probe
set -x
mkdir -p gpurun_out
GCNB_TEST_UNVERIFIED=1 timeout 400 python -m pytest tests/test_gpu_optin.py -m gpu -q > gpurun_out/r2_pytest_optin.log 2>&1
timeout 90 python tools/pdl_probe.py cbg 40 > gpurun_out/r2_pdl_probe_cbg.txt 2>&1
timeout 90 python tools/variant_sweep.py 13,16,14,17 20 > gpurun_out/r2_variant_sweep_persistent.txt 2>&1
timeout 120 python tools/bn_probe.py 20 > gpurun_out/r2_bn_probe.txt 2>&1
timeout 400 python tools/rmat_probe.py --widths 256,100,48 --relabel none,degree,random --bf16 --reps 3 > gpurun_out/r2_rmat_probe.txt 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,lts__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active
timeout 400 ncu --metrics $M -k regex:spmm_ --clock-control none --csv --log-file gpurun_out/r2_ncu_rmat_probe.csv \
  python tools/rmat_probe.py --widths 256,100 --relabel none,degree --reps 1 > gpurun_out/r2_ncu_rmat_probe.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_smi.txt
