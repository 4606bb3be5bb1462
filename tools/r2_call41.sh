#!/usr/bin/env bash
# on the final defaults: 5 / 6 node steps, 3 / 4 steps in the closest-hit kernel; run-time knobs once more
set -u
O=gpurun_out
export SWEEP_LIBS="libdsrt.so,libdsrt_ns5.so,libdsrt_ns6.so,libdsrt_cs3.so,libdsrt_cs4.so,libdsrt.so" SWEEP_OPTS='[{}]'
SWEEP_SCENE=c2 python tools/sweeps/sweep_variants.py 64 > $O/r2c41_c2.log 2>&1; cat $O/r2c41_c2.log
SWEEP_SCENE=c3 python tools/sweeps/sweep_variants.py 64 > $O/r2c41_c3.log 2>&1; cat $O/r2c41_c3.log
export SWEEP_LIBS="libdsrt.so,libdsrt_ns6.so,libdsrt_cs3.so,libdsrt_cs4.so"
SWEEP_SCENE=soup8 python tools/sweeps/sweep_variants.py 8 > $O/r2c41_soup8.log 2>&1; cat $O/r2c41_soup8.log
export SWEEP_LIBS="libdsrt.so" SWEEP_OPTS='[{}, {"postpone_min_lanes": 6}, {"postpone_min_lanes": 12}, {"postpone_min_lanes": 16}, {"coop_min_pairs": 10}, {"refill_busy_lanes": 16}, {"refill_busy_lanes": 20}, {"postpone_wait_mode": 1}, {"refill_patience": 4}, {"refill_patience": 10}, {}]'
SWEEP_SCENE=c2 python tools/sweeps/sweep_variants.py 64 > $O/r2c41_knobs_c2.log 2>&1; cat $O/r2c41_knobs_c2.log
