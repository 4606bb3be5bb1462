#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
L="libdsrt.so,libdsrt_cs1.so,libdsrt_cs3.so,libdsrt_cs4.so,libdsrt.so"
SWEEP_SCENE=c2 SWEEP_LIBS=$L SWEEP_OPTS='[{}]' python tools/sweeps/sweep_variants.py 64 > $O/r2c16_sweep_c2.log 2>&1; cat $O/r2c16_sweep_c2.log
SWEEP_SCENE=soup8 SWEEP_LIBS=$L SWEEP_OPTS='[{}, {"postpone_min_lanes": 4}, {"postpone_min_lanes": 16}, {"refill_busy_lanes": 24}]' python tools/sweeps/sweep_variants.py 8 > $O/r2c16_sweep_soup8.log 2>&1; cat $O/r2c16_sweep_soup8.log
