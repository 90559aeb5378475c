#!/bin/bash
# Weak-scaling sweep, second pass (arenas tiled around the origin): c2 at N = 2, 4, 8 and c4 at N = 8.
TAG=${1:-r3s}
O=gpurun_out
mkdir -p $O
for N in ${NS:-2 4 8}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29620+N)) bench.py --gpus $N > $O/${TAG}_c2_${N}gpu.json 2> $O/${TAG}_c2_${N}gpu.err
  python -c "
import json;d=json.load(open('$O/${TAG}_c2_${N}gpu.json'));print('c2 N=$N value %.0f e2e %.0f ms/step %.2f ok %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['replans_ok_frac']), d['own_ms_per_step_by_rank'])"
done
if [ -z "$NO_C4" ]; then
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29640 bench.py --gpus 8 --config c4 --steps 10 > $O/${TAG}_c4_8gpu.json 2> $O/${TAG}_c4_8gpu.err
python -c "
import json;d=json.load(open('$O/${TAG}_c4_8gpu.json'));print('c4 N=8 value %.0f e2e %.0f ms/step %.2f ok %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['replans_ok_frac']), d['config']['agents_total'], d['own_ms_per_step_by_rank'], d['stage_ms_serialised'])"
fi
