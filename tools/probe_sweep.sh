set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 0 1 2 3 4; do GCNB_SPMM_GROUP_VARIANT=$v python tools/spmm_probe.py cbg 32 --check; done
for v in 0 2; do GCNB_SPMM_GROUP_VARIANT=$v python tools/spmm_probe.py 400000:25 32 --check; done
for v in 0 2; do GCNB_SPMM_GROUP_VARIANT=$v python tools/spmm_probe.py 100000:100 64 --check; done
for v in 0 2; do GCNB_SPMM_GROUP_VARIANT=$v python tools/spmm_probe.py 100000:100 16 --check; done
python tools/spmm_probe.py 100000:100 24 --check
python tools/spmm_probe.py 100000:100 8 --check
