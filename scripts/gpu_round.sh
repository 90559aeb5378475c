#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, the ncu launch list of the same command and one
# `ncu --set full` capture of the voxel-update kernels.  Usage: scripts/gpu_round.sh TAG [skip-tests]
TAG=${1:-r1x}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $O/${TAG}_pytest.log
  tail -3 $O/${TAG}_pytest.log
fi
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1
echo "smoke exit $?" >> $O/${TAG}_smoke.log
tail -2 $O/${TAG}_smoke.log
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "bench exit $?"
cat $O/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vox -s 8 -c 4 \
  -o $O/${TAG}_voxel -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_voxel.log 2>&1
echo "ncu voxel exit $?"
ls -la $O | tail -20
