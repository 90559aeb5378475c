#!/bin/bash
# ncu --set full capture (with source) of the whole-batch A* launch of scripts/dev_stages.py.  Usage: scripts/gpu_astar_ncu.sh TAG
TAG=${1:-r2n}
O=gpurun_out
mkdir -p $O
POP_STEPS=8 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_astar -s 8 -c 1 \
  -o $O/${TAG}_k_astar -f python scripts/dev_stages.py > $O/${TAG}_ncu_k_astar.log 2>&1
echo "ncu exit $?"; tail -3 $O/${TAG}_ncu_k_astar.log; ls -la $O/${TAG}_k_astar.ncu-rep
