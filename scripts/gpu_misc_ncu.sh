#!/bin/bash
# ncu --set full captures of the kernels the round-end pass does not reach: raycast map, depth sensor, peer projection,
# separation.  Raw pages are exported on the box (gpurun_out/ is size-limited).  Usage: scripts/gpu_misc_ncu.sh TAG
TAG=${1:-r3m}
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --clock-control none -k regex:"k_gm_(points|rays|update|clear|inflate|project|frame_reset)" -s 14 -c 14 \
  -o /tmp/${TAG}_gm -f python scripts/gridmap_bench.py 64 6 > $O/${TAG}_ncu_gm.log 2>&1
echo "ncu gm exit $?"; ncu -i /tmp/${TAG}_gm.ncu-rep --page raw --csv > $O/${TAG}_gm_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:"k_depth" -c 6 \
  -o /tmp/${TAG}_depth -f python -m pytest tests/test_gpu_render.py -x -q > $O/${TAG}_ncu_depth.log 2>&1
echo "ncu depth exit $?"; ncu -i /tmp/${TAG}_depth.ncu-rep --page raw --csv > $O/${TAG}_depth_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:"k_peer|k_separate" -s 40 -c 8 \
  -o /tmp/${TAG}_peer -f python bench.py --steps 2 --warmup 3 --no-dsp --no-cpu-baseline > $O/${TAG}_ncu_peer.log 2>&1
echo "ncu peer/separate exit $?"; ncu -i /tmp/${TAG}_peer.ncu-rep --page raw --csv > $O/${TAG}_peer_raw.csv 2>/dev/null
ls -la $O/${TAG}_*raw.csv | awk '{print $5, $9}'
