#!/bin/bash
# One gpurun call: ncu --set full captures (with source) of the planning-stage kernels inside the bench
# workload, plus the stage-by-stage statistics of scripts/dev_stages.py.  Usage: scripts/gpu_profile_r2.sh TAG
TAG=${1:-r2a}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 300 python scripts/dev_stages.py > $O/${TAG}_stages.txt 2>&1
echo "stages exit $?"; tail -30 $O/${TAG}_stages.txt
for K in k_qp_admm k_corridor_firi k_separate k_astar k_corridor_points; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 3 \
    -o $O/${TAG}_$K -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dsp > $O/${TAG}_ncu_$K.log 2>&1
  echo "ncu $K exit $?"
done
ls -la $O | tail -20
