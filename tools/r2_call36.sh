#!/usr/bin/env bash
# regroup_top makes half of the shadow rays one visit shorter: does it pay with an earlier refill?  regroup x refill thresholds on c2 / c3
set -u
O=gpurun_out
export SWEEP_LIBS="libdsrt.so" SWEEP_OPTS='[{}, {"regroup_top": 1}, {"regroup_top": 1, "refill_busy_lanes": 20}, {"regroup_top": 1, "refill_busy_lanes": 22}, {"regroup_top": 1, "refill_busy_lanes": 24}, {"regroup_top": 1, "refill_busy_lanes": 20, "refill_hi_lanes": 28}, {"regroup_top": 1, "refill_patience": 3}, {"regroup_top": 1, "postpone_min_lanes": 6}, {"regroup_top": 1, "postpone_wait_mode": 1}, {}]'
SWEEP_SCENE=c2 python tools/sweeps/sweep_variants.py 64 > $O/r2c36_c2.log 2>&1; cat $O/r2c36_c2.log
SWEEP_SCENE=c3 python tools/sweeps/sweep_variants.py 64 > $O/r2c36_c3.log 2>&1; cat $O/r2c36_c3.log
