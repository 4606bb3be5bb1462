#!/usr/bin/env bash
# ordered vs unordered any-hit children on the soups (long rays, unbounded shadow rays)
set -u
O=gpurun_out
export SWEEP_LIBS="libdsrt.so,libdsrt_unord.so,libdsrt_ns4u.so,libdsrt.so,libdsrt_unord.so" SWEEP_OPTS='[{}]'
SWEEP_SCENE=soup1 python tools/sweeps/sweep_variants.py 8 > $O/r2c39_soup1.log 2>&1; cat $O/r2c39_soup1.log
export SWEEP_LIBS="libdsrt.so,libdsrt_unord.so,libdsrt_ns4u.so"
SWEEP_SCENE=soup8 python tools/sweeps/sweep_variants.py 8 > $O/r2c39_soup8.log 2>&1; cat $O/r2c39_soup8.log
