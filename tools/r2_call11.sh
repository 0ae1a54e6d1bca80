#!/bin/bash
# round 2, call 11: (a) TMA gather4 microbenchmark, (b) column-sliced SpMM through the existing kernels (+ DRAM bytes by ncu)
set -x
mkdir -p gpurun_out
G=tools/microbench/gather4_bw
{
for cfg in "48000 256 256 8 6 0 0" "48000 256 256 8 6 0 0 4" "2400000 256 256 8 6 0 0" "2400000 256 256 8 6 48000 50" \
           "2400000 256 256 4 12 48000 50" "2400000 256 256 16 3 48000 50" "2400000 128 256 8 12 96000 62" \
           "2400000 64 256 16 12 192000 75" "48000 64 256 16 12 0 0" "48000 128 128 8 12 0 0" "2400000 100 100 8 12 120000 65"; do
  timeout 120 $G $cfg
done
} > gpurun_out/r2_gather4_bw.txt 2>&1
timeout 600 python tools/rmat_probe.py --widths 256 --cpu-edges --check --reps 4 --slices 1,2,4,8 \
   --sweep 1:48:0:0,0:0:0:0,2:48:0:8 > gpurun_out/r2_rmat_slices.txt 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
   --clock-control none -k regex:spmm --csv --log-file gpurun_out/r2_ncu_rmat_slices.csv \
   python tools/rmat_probe.py --widths 256 --cpu-edges --reps 1 --fwd-only --slices 1,2,4,8 --sweep 1:48:0:0 > gpurun_out/r2_ncu_rmat_slices.log 2>&1
tail -3 gpurun_out/r2_gather4_bw.txt; grep "f=256\|slices\|config" gpurun_out/r2_rmat_slices.txt | tail -40
