#!/bin/bash
# round 2, final single-GPU validation: the full GPU suite, smoke(), the default bench line and the reference arm
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.txt 2>&1; tail -3 gpurun_out/r2_smoke_final.txt
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; tail -c 600 gpurun_out/r2_bench_final.err; head -c 400 gpurun_out/r2_bench_final.json
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_reference_final.json 2> gpurun_out/r2_bench_reference_final.err; head -c 600 gpurun_out/r2_bench_reference_final.json
