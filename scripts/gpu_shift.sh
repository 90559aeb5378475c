#!/bin/bash
# Diagnosis: one GPU flying a later rank's (translated) arena; step and stage times per shift.
O=gpurun_out; mkdir -p $O
for S in 0 1 3 7; do
  POP_BENCH_ARENA_SHIFT=$S timeout 600 python bench.py --steps 10 --warmup 3 --no-dsp --no-cpu-baseline > $O/r2t_shift$S.json 2> $O/r2t_shift$S.err
  python -c "
import json;d=json.load(open('$O/r2t_shift$S.json'));print('shift $S', round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['stage_ms_serialised'].items()}, d['replans_ok_frac'])"
done
