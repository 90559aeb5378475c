#!/bin/bash
# Rebuilds k_qp.cu with -DPOP_QP_PHASE_PROF=1 (clock64 around the phases of an ADMM iteration) on the GPU box, in the
# box's scratch copy only, and prints the per-phase cycles of the bench workload's QPs.  Usage: scripts/gpu_qp_phase.sh TAG
TAG=${1:-r2b}
O=gpurun_out
mkdir -p $O
cd pred_occ_planner_b200/csrc
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC --expt-relaxed-constexpr \
  -DPOP_QP_PHASE_PROF=1 -c k_qp.cu -o k_qp.o 2> ../../$O/${TAG}_ptxas.log && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libpop.so *.o -lcudart
cd ../..
timeout 300 python scripts/dev_stages.py > $O/${TAG}_stages_phase.txt 2>&1
tail -16 $O/${TAG}_stages_phase.txt
