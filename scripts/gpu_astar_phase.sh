#!/bin/bash
# Rebuilds k_astar.cu with -DPOP_ASTAR_PROF (clock64 around the sections of one expansion) on the GPU box, in the box's
# scratch copy only, and prints the per-section cycles of the bench workload's searches.  Usage: scripts/gpu_astar_phase.sh TAG
TAG=${1:-r2u}
O=gpurun_out
mkdir -p $O
cd pred_occ_planner_b200/csrc
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -ffp-contract=off --expt-relaxed-constexpr \
  -DPOP_ASTAR_PROF=1 -Xptxas -v -c k_astar.cu -o k_astar.o 2> ../../$O/${TAG}_ptxas.log && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libpop.so *.o -lcudart -ldl
cd ../..
POP_STEPS=${POP_STEPS:-8} timeout 600 python scripts/dev_stages.py > $O/${TAG}_astar_phase.txt 2>&1
grep -a astar_prof $O/${TAG}_astar_phase.txt | sort -t' ' -k5 -n -r | head -30
tail -12 $O/${TAG}_astar_phase.txt
