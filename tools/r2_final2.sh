#!/bin/bash
# round 2: after the bench's config / run split: one N=1 line and one N=2 line (both must print and carry the same config object)
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err; tail -c 300 gpurun_out/r2_bench_final_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_final_n2.json 2> gpurun_out/r2_bench_final_n2.err; grep -v Warn gpurun_out/r2_bench_final_n2.err | tail -5 | cut -c1-300
python - <<'PY'
import json
a=json.loads([l for l in open('gpurun_out/r2_bench_final_n1.json') if l.startswith('{')][-1])
b=json.loads([l for l in open('gpurun_out/r2_bench_final_n2.json') if l.startswith('{')][-1])
print('n1', a['ms_per_step'], a['parity']['ok'], sorted(a['run']))
print('n2', b['ms_per_step'], b['parity']['ok'], sorted(b['run']))
print('same config', a['config']==b['config'])
PY
