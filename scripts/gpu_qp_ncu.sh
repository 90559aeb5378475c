#!/bin/bash
# QP parity tests + one ncu --set full capture (with source) of the batched QP kernel of scripts/dev_stages.py.
TAG=${1:-r2g}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_qp.py tests/test_gpu_cycle.py -x -q 2>&1 | tail -4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_qp_admm -c 1 -o $O/${TAG}_qp -f python scripts/dev_stages.py > $O/${TAG}_ncu.log 2>&1
echo "ncu exit $?"; grep "qp status" $O/${TAG}_ncu.log
