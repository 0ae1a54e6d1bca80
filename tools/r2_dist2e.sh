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
    print('$name', 'ms', round(d['ms_per_step'],2), 'split', d.get('run', d['config'])['row_block_split'], d.get('run', d['config'])['partition'][:40], 'exch+spmm', round(d['roofline']['exchange_plus_spmm_ms'],2), 'spmm', round(d['roofline']['kernel_ms'],2), 'graph', d.get('run', d['config'])['cuda_graph'], 'halo', d.get('run', d['config'])['halo'])
except Exception as e:
    print('$name failed', e)
    print(open('gpurun_out/r2_tune_n${N}_$name.err').read()[-1500:])
PY
}
run bal_unsplit GCNB_DIST_SPLIT=0
run bal_split GCNB_DIST_SPLIT=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_products_n${N}.json 2> gpurun_out/r2_bench_products_n${N}.err
grep -v Warn gpurun_out/r2_bench_products_n${N}.err | tail -5 | cut -c1-300; head -c 400 gpurun_out/r2_bench_products_n${N}.json
