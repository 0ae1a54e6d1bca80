run() { timeout 120 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 tools/nccl_ag_probe.py 2>&1 | grep -E "all-gather|rror" | head -3; }
run X=1
run NCCL_ALGO=Ring
run NCCL_ALGO=NVLS
run NCCL_MIN_NCHANNELS=32
run NCCL_PROTO=Simple NCCL_MIN_NCHANNELS=24
run NCCL_NVLS_ENABLE=0
