#!/usr/bin/env bash
# (1) is the regroup_top loss on c2 the extra stack level (shared memory)?  stack slack 1 / 0 x regroup 0 / 1 on c2 and c3
# (2) run-time knobs re-swept after the primitive tests per segment fell by 37 %
set -u
O=gpurun_out
export SWEEP_LIBS="libdsrt.so,libdsrt_s0.so,libdsrt.so,libdsrt_s0.so" SWEEP_OPTS='[{}, {"regroup_top": 1}]'
SWEEP_SCENE=c2 python tools/sweeps/sweep_variants.py 64 > $O/r2c33_slack_c2.log 2>&1; cat $O/r2c33_slack_c2.log
export SWEEP_LIBS="libdsrt.so,libdsrt_s0.so"
SWEEP_SCENE=c3 python tools/sweeps/sweep_variants.py 64 > $O/r2c33_slack_c3.log 2>&1; cat $O/r2c33_slack_c3.log
export SWEEP_LIBS="libdsrt.so" SWEEP_OPTS='[{}, {"postpone_min_lanes": 6}, {"postpone_min_lanes": 10}, {"postpone_min_lanes": 12}, {"coop_min_pairs": 4}, {"coop_min_pairs": 10}, {"coop_min_pairs": 16}, {"refill_busy_lanes": 16}, {"refill_busy_lanes": 20}, {"refill_busy_lanes": 22}, {"postpone_wait_mode": 1}, {"postpone_wait_mode": 4}, {"refill_patience": 4}, {"refill_hi_lanes": 24}, {}]'
SWEEP_SCENE=c2 python tools/sweeps/sweep_variants.py 64 > $O/r2c33_knobs_c2.log 2>&1; cat $O/r2c33_knobs_c2.log
