#!/bin/bash
# round 2: the pipelined (column-chunked) halo exchange + in-place own slot at N GPUs: parity (dist_check), then timing
set -x
mkdir -p gpurun_out
N=${1:-2}
GCNB_DIST_CHECK_EXCHANGES=nccl,halo timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  tools/dist_check.py > gpurun_out/r2_dist_check_${N}gpu_inplace.txt 2>&1
grep "exchange=\|ALL OK\|MISMATCH\|Error" gpurun_out/r2_dist_check_${N}gpu_inplace.txt | cut -c1-250
GCNB_DIST_CHUNK_MIN_COLS=16 GCNB_DIST_CHECK_EXCHANGES=halo timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  tools/dist_check.py > gpurun_out/r2_dist_check_${N}gpu_chunked.txt 2>&1
grep "exchange=\|ALL OK\|MISMATCH\|Error" gpurun_out/r2_dist_check_${N}gpu_chunked.txt | cut -c1-250
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus $N --steps 8 --warmup 3 --quick > gpurun_out/r2_tune_n${N}_$name.json 2> gpurun_out/r2_tune_n${N}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_tune_n${N}_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms', round(d['ms_per_step'],2), 'split', d.get('run', d['config'])['row_block_split'], 'exch+spmm', round(d['roofline']['exchange_plus_spmm_ms'],2), 'spmm', round(d['roofline']['kernel_ms'],2), 'graph', d.get('run', d['config'])['cuda_graph'])
except Exception as e:
    print('$name failed', e)
    print(open('gpurun_out/r2_tune_n${N}_$name.err').read()[-1500:])
PY
}
run chunks2 GCNB_DIST_CHUNKS=2
run chunks1 GCNB_DIST_CHUNKS=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_products_n${N}.json 2> gpurun_out/r2_bench_products_n${N}.err
grep -v Warn gpurun_out/r2_bench_products_n${N}.err | tail -5 | cut -c1-300; head -c 600 gpurun_out/r2_bench_products_n${N}.json
