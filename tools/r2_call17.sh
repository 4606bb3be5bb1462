#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
OPTS='[{"refill_busy_lanes": 18}, {"refill_busy_lanes": 20}, {"refill_busy_lanes": 22}, {"refill_busy_lanes": 24}, {"refill_busy_lanes": 26}, {"refill_busy_lanes": 28}, {"refill_busy_lanes": 30}, {"refill_busy_lanes": 24, "postpone_min_lanes": 4}, {"refill_busy_lanes": 28, "postpone_min_lanes": 4}]'
for sc in c2 soup1 soup8; do
  spp=8; [ $sc = c2 ] && spp=64
  SWEEP_SCENE=$sc SWEEP_LIBS=libdsrt.so SWEEP_OPTS="$OPTS" python tools/sweeps/sweep_variants.py $spp > $O/r2c17_sweep_$sc.log 2>&1; echo "== $sc"; cat $O/r2c17_sweep_$sc.log
done
SWEEP_SCENE=soup64 SWEEP_DEVICE_BUILD=1 SWEEP_LIBS=libdsrt.so SWEEP_OPTS="$OPTS" python tools/sweeps/sweep_variants.py 8 > $O/r2c17_sweep_soup64.log 2>&1; echo "== soup64"; cat $O/r2c17_sweep_soup64.log
