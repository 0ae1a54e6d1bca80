#!/bin/bash
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
    print('$name', 'ms', round(d['ms_per_step'],2), 'split', d['config']['row_block_split'], d['config']['partition'], 'exch+spmm', round(d['roofline']['exchange_plus_spmm_ms'],2), 'spmm', round(d['roofline']['kernel_ms'],2), 'bounds', d['config']['bounds'], 'graph', d['config']['cuda_graph'])
except Exception as e:
    print('$name failed', e)
PY
}
run split_w12 GCNB_DIST_SPLIT=1 GCNB_DIST_ROW_WEIGHT=12
run unsplit_w12 GCNB_DIST_SPLIT=0 GCNB_DIST_ROW_WEIGHT=12
run unsplit_w25 GCNB_DIST_SPLIT=0 GCNB_DIST_ROW_WEIGHT=25
run unsplit_w6 GCNB_DIST_SPLIT=0 GCNB_DIST_ROW_WEIGHT=6
run unsplit_w12_nccl GCNB_DIST_SPLIT=0 GCNB_DIST_ROW_WEIGHT=12 GCNB_DIST_EXCHANGE=nccl
