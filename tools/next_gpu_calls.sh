#!/bin/bash
# The measurements prepared at the end of round 1 (GPU budget spent), in the order they should be run.
# Each block is one gpurun call; copy what matters from gpurun_out/ into profiles/.
#
# 1 GPU (about a minute of box time):
#   gpurun --timeout 420 -- 'bash tools/next_gpu_calls.sh one'
# 2 GPUs (charged 2x):
#   gpurun --gpus 2 --timeout 300 -- 'bash tools/next_gpu_calls.sh two'
# 8 GPUs (charged 8x; about a minute):
#   gpurun --gpus 8 --timeout 300 -- 'bash tools/next_gpu_calls.sh eight'
set -x
mkdir -p gpurun_out
case "$1" in
one)
  # the opt-in paths' own parity tests (skipped by default because they had never run on hardware)
  GCNB_TEST_UNVERIFIED=1 timeout 240 python -m pytest tests/test_gpu_optin.py -m gpu -x -q > gpurun_out/pytest_optin.log 2>&1
  # programmatic dependent launch of the step's kernel chain: bit-identity + graph-replay time, off / on
  timeout 60 python tools/pdl_probe.py cbg 40 > gpurun_out/pdl_probe_cbg.txt 2>&1
  # the persistent group kernel with the next row set prefetched (variants 16 / 17) against the defaults (13 / 14)
  timeout 60 python tools/variant_sweep.py 13,16,14,17 20 > gpurun_out/variant_sweep_persistent.txt 2>&1
  # the bf16 panel kernels were not in the CTA-shape sweeps: 64-byte rows stay on variant 0 until this says otherwise
  timeout 60 python tools/variant_sweep.py 0,2,13,14 20 --bf16 > gpurun_out/variant_sweep_bf16.txt 2>&1
  # bare-gather floor again, now with mode G: the gathers as 16-byte cp.async into shared memory (lines in flight bounded
  # by shared memory instead of registers) -- decides whether a cp.async-gather SpMM variant is worth writing
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench/gather_bw tools/microbench/gather_bw.cu \
    && timeout 90 tools/microbench/gather_bw > gpurun_out/gather_bw_cpasync.txt 2>&1
  # (ReLU ->) fresh BatchNorm: parity and time against torch's relu + batch_norm
  timeout 90 python tools/bn_probe.py 20 > gpurun_out/bn_probe.txt 2>&1
  # with PDL on, the whole bench line
  GCNB_PDL=1 timeout 100 python bench.py --no-cpu-baseline > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err
  ;;
two)
  # parity of every exchange with the single-GPU layer at 2 GPUs, eager and graph replay; "nvls" = the multicast
  # exchange (own multimem.st push into torch symmetric memory), written but never run in round 1
  GCNB_DIST_CHECK_EXCHANGES=peer,nccl,nccl-bf16,nvls,halo timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
    --master-addr 127.0.0.1 --master-port 29513 tools/dist_check.py > gpurun_out/dist_check_2gpu_nvls.txt 2>&1
  # the exact-size slot exchange (grouped send / recv instead of the padded all-gather) and the column-chunked one
  GCNB_DIST_EXACT_SLOTS=1 GCNB_DIST_CHECK_EXCHANGES=nccl timeout 120 python -m torch.distributed.run --nnodes=1 \
    --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/dist_check.py > gpurun_out/dist_check_2gpu_exact.txt 2>&1
  GCNB_DIST_NCCL_CHUNKS=2 GCNB_DIST_CHECK_EXCHANGES=nccl timeout 120 python -m torch.distributed.run --nnodes=1 \
    --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/dist_check.py > gpurun_out/dist_check_2gpu_chunks.txt 2>&1
  ;;
eight)
  GCNB_DIST_EXCHANGE=nvls timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
    --master-addr 127.0.0.1 --master-port 29509 bench.py --gpus 8 --steps 20 --warmup 5 \
    > gpurun_out/bench_n8_nvls.json 2> gpurun_out/bench_n8_nvls.err
  # the double-buffered end-to-end loop of the multi-GPU arm (opt-in until this has run)
  GCNB_BENCH_E2E=pipelined timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
    --master-addr 127.0.0.1 --master-port 29510 bench.py --gpus 8 --steps 20 --warmup 5 \
    > gpurun_out/bench_n8_e2e_pipelined.json 2> gpurun_out/bench_n8_e2e_pipelined.err
  # the bf16 panel tier of the row-partitioned layer beside the fp32 step (half the exchange bytes)
  GCNB_BENCH_DIST_BF16=1 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
    --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 8 --steps 20 --warmup 5 \
    > gpurun_out/bench_n8_bf16_tier.json 2> gpurun_out/bench_n8_bf16_tier.err
  for chunks in 1 2; do
    GCNB_DIST_NCCL_CHUNKS=$chunks timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
      --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 \
      > gpurun_out/bench_n8_chunks$chunks.json 2> gpurun_out/bench_n8_chunks$chunks.err
  done
  ;;
papers) ;;  # below
*)
  echo "usage: $0 one|two|eight|papers"; exit 2;;
esac
# BASELINE configs[4], never run end to end in round 1 (one rank's share was, tools/papers_block.py):
#   gpurun --gpus 8 --timeout 600 -- 'bash tools/next_gpu_calls.sh papers'
if [ "$1" = papers ]; then
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 8 --workload papers --steps 5 --warmup 3 > gpurun_out/bench_n8_papers.json 2> gpurun_out/bench_n8_papers.err
fi
