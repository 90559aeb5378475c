#!/bin/bash
# two-rank exchange test (strict + pipelined) and c2 at N = 2 on a 2-GPU box
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_exchange.py tests/test_gpu_astar.py -x -q 2>&1 | grep -a "passed\|failed\|rror\|assert" | tail -5
for MODE in pipelined strict; do
POP_BENCH_EXCHANGE=$MODE timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29720 bench.py --gpus 2 > $O/r3t_c2_2gpu_$MODE.json 2> $O/r3t_c2_2gpu_$MODE.err
python -c "
import json;d=json.load(open('$O/r3t_c2_2gpu_$MODE.json'));print('$MODE c2 N=2 value %.0f e2e %.0f ms/step %.2f ok %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['replans_ok_frac']), d['own_ms_per_step_by_rank'])"
done
