#!/bin/bash
# One QP tuning iteration on the GPU box: parity tests of the QP and the cycle, then the per-phase cycles
# (k_qp.cu rebuilt with -DPOP_QP_PHASE_PROF=1 in the box's scratch copy).  Usage: scripts/gpu_qp_iter.sh TAG
TAG=${1:-r2q}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_qp.py tests/test_gpu_cycle.py -x -q 2>&1 | tail -6
timeout 300 python scripts/dev_stages.py > $O/${TAG}_stages.txt 2>&1; grep "qp status\|ms " $O/${TAG}_stages.txt | head -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-dsp > $O/${TAG}_bench.json 2>$O/${TAG}_bench.err; python -c "
import json;d=json.load(open('$O/${TAG}_bench.json'));print('ms_per_step',d['ms_per_step'],'value',d['value'],'stages',d.get('stage_ms_serialised'))"
bash scripts/gpu_qp_phase.sh $TAG | tail -14
