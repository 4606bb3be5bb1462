#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
OPTS='[{"refill_patience": 1000000}, {"refill_patience": 4}, {"refill_patience": 6}, {"refill_patience": 8}, {"refill_patience": 10}, {"refill_patience": 14}, {"refill_patience": 20}, {"refill_patience": 8, "refill_hi_lanes": 24}, {"refill_patience": 8, "refill_hi_lanes": 28}, {"refill_patience": 12, "refill_hi_lanes": 24}]'
for sc in c2 soup1 soup8; do
  spp=8; [ $sc = c2 ] && spp=64
  SWEEP_SCENE=$sc SWEEP_LIBS=libdsrt.so SWEEP_OPTS="$OPTS" python tools/sweeps/sweep_variants.py $spp > $O/r2c18_sweep_$sc.log 2>&1; echo "== $sc"; cut -c20-170 $O/r2c18_sweep_$sc.log
done
