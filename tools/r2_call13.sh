#!/bin/bash
# round 2, call 13: are hub rows pinned in L2 (persisting access-policy window) worth anything for the gather?
mkdir -p gpurun_out
G=tools/microbench/gather4_bw
{
for cfg in "2400000 256 256 8 6 60000 60 1 0" "2400000 256 256 8 6 60000 60 1 64" "2400000 256 256 8 6 60000 60 1 96" \
           "2400000 256 256 8 6 32000 50 1 0" "2400000 256 256 8 6 32000 50 1 40" "2400000 256 256 8 6 100000 69 1 0" "2400000 256 256 8 6 100000 69 1 110"; do
  timeout 120 $G $cfg
done
} > gpurun_out/r2_gather_persist.txt 2>&1
grep -v "^gather4\|check" gpurun_out/r2_gather_persist.txt
