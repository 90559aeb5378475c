#!/bin/bash
# Round-end GPU pass in one gpurun call: parity suite, smoke, both bench arms, the ncu launch list of the bench command,
# ncu --set full captures of the voxel update, A*, the particle map and the raycast map kernels.  Usage: scripts/gpu_final.sh TAG
TAG=${1:-r3f}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/${TAG}_gpu.txt 2>&1
nproc >> $O/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $O/${TAG}_pytest.log; grep -a "passed\|failed\|rror\|exit" $O/${TAG}_pytest.log | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -1 $O/${TAG}_smoke.log
timeout 600 python bench.py --impl reference > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; echo "ref exit $?"; cut -c1-300 $O/${TAG}_bench_ref.json
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench exit $?"; cat $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv \
  --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_vox -s 8 -c 4 \
  -o $O/${TAG}_voxel -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dsp > $O/${TAG}_ncu_voxel.log 2>&1
echo "ncu voxel exit $?"
POP_STEPS=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_astar -s 9 -c 1 \
  -o $O/${TAG}_k_astar -f python scripts/dev_stages.py > $O/${TAG}_ncu_astar.log 2>&1
echo "ncu astar exit $?"
# particle map and raycast map: the raw pages are exported here and the reports dropped (gpurun_out/ is size-limited)
timeout 600 ncu --set full --clock-control none -k regex:k_dsp -s 112 -c 28 \
  -o /tmp/${TAG}_dsp -f python scripts/dsp_bench.py --agents 64 --frames 6 --big-agents 0 > $O/${TAG}_ncu_dsp.log 2>&1
echo "ncu dsp exit $?"
ncu -i /tmp/${TAG}_dsp.ncu-rep --page raw --csv > $O/${TAG}_dsp_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:"k_gm_|k_depth" -s 24 -c 8 \
  -o /tmp/${TAG}_gm -f python scripts/gridmap_bench.py 64 6 > $O/${TAG}_ncu_gm.log 2>&1
echo "ncu gm exit $?"
ncu -i /tmp/${TAG}_gm.ncu-rep --page raw --csv > $O/${TAG}_gm_raw.csv 2>/dev/null
ls -la $O/${TAG}_* | awk '{print $5, $9}'
