#!/bin/bash
# step time vs number of agent groups the planning stages are pipelined over (POP_PLAN_GROUPS), both arena placements
O=gpurun_out
for S in 0 1; do for G in 1 4 8 16 32 64; do
  POP_BENCH_ARENA_SHIFT=$S POP_PLAN_GROUPS=$G python bench.py --steps 10 --warmup 3 --no-dsp --no-cpu-baseline > $O/grp_${S}_${G}.json 2>$O/grp.err
  python -c "
import json;d=json.load(open('$O/grp_${S}_${G}.json'));print('shift',$S,'groups',$G,'ms/step %.2f'%d['ms_per_step'],'e2e %.2f'%d['e2e']['ms_per_step'],'launches',d['gpu_launches'])"
done; done
