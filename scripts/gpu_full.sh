#!/bin/bash
# Full GPU check: parity suite, smoke, both bench arms.  Usage: scripts/gpu_full.sh TAG
TAG=${1:-r2x}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log; tail -4 $O/${TAG}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/${TAG}_smoke.log
timeout 600 python bench.py --impl reference > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "ref exit $?"; cat $O/${TAG}_bench_ref.json | cut -c1-400
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"; cat $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
