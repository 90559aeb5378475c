#!/bin/bash
# Multi-GPU check on an N-GPU box: exchange test + bench at N ranks.  Usage: scripts/gpu_multi.sh TAG N [config]
TAG=${1:-r2m}; N=${2:-2}; CFG=${3:-c2}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name --format=csv,noheader | head -8 > $O/${TAG}_gpus.txt
timeout 600 python -m pytest tests/test_gpu_exchange.py -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config $CFG --steps 20 --warmup 3 > $O/${TAG}_bench_${N}gpu.json 2> $O/${TAG}_bench_${N}gpu.err; echo "bench exit $?"; cut -c1-700 $O/${TAG}_bench_${N}gpu.json; tail -3 $O/${TAG}_bench_${N}gpu.err
