#!/bin/bash
# Weak-scaling sweep on an 8-GPU box: c2 at N = 1, 2, 4, 8 and c4 (1 024 agents) at N = 8; lines kept for profiles/.
TAG=${1:-r2s}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name --format=csv,noheader | head -8 > $O/${TAG}_gpus.txt; nproc >> $O/${TAG}_gpus.txt
timeout 600 python -m pytest tests/test_gpu_exchange.py -x -q 2>&1 | tail -3
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    timeout 900 python bench.py --no-dsp > $O/${TAG}_c2_${N}gpu.json 2> $O/${TAG}_c2_${N}gpu.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N > $O/${TAG}_c2_${N}gpu.json 2> $O/${TAG}_c2_${N}gpu.err
  fi
  python -c "
import json;d=json.load(open('$O/${TAG}_c2_${N}gpu.json'));print('c2 N=$N value %.0f e2e %.0f ms/step %.2f ok %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['replans_ok_frac']))"
done
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --config c4 --steps 10 > $O/${TAG}_c4_8gpu.json 2> $O/${TAG}_c4_8gpu.err
python -c "
import json;d=json.load(open('$O/${TAG}_c4_8gpu.json'));print('c4 N=8 value %.0f e2e %.0f ms/step %.2f ok %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['replans_ok_frac']), d['config']['agents_total'], d['stage_ms_serialised'])"
timeout 600 python bench.py --impl reference --config c4 --steps 3 --warmup 1 > $O/${TAG}_c4_ref.json 2> $O/${TAG}_c4_ref.err; cut -c1-300 $O/${TAG}_c4_ref.json
