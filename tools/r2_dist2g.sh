#!/bin/bash
# round 2: high-priority side stream for the exchange; chunked / unchunked / split at N GPUs
set -x
mkdir -p gpurun_out
N=${1:-2}
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus $N --steps 8 --warmup 3 --quick > gpurun_out/r2_tune_n${N}_$name.json 2> gpurun_out/r2_tune_n${N}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_tune_n${N}_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms', round(d['ms_per_step'],2), 'split', d.get('run', d['config'])['row_block_split'], 'exch+spmm', round(d['roofline']['exchange_plus_spmm_ms'],2), 'spmm', round(d['roofline']['kernel_ms'],2), 'chunks', d['roofline'].get('exchange_column_chunks'))
except Exception as e:
    print('$name failed', e)
    print(open('gpurun_out/r2_tune_n${N}_$name.err').read()[-1500:])
PY
}
run prio_chunks2 GCNB_DIST_CHUNKS=2
run prio_chunks1 GCNB_DIST_CHUNKS=1
run prio_split GCNB_DIST_SPLIT=1
run noprio_chunks2 GCNB_DIST_CHUNKS=2 GCNB_DIST_SIDE_PRIORITY=0
