#!/usr/bin/env bash
# unordered any-hit children combined with 3 / 4 node steps per primitive decision (and 3 steps in the closest-hit kernel)
set -u
O=gpurun_out
export SWEEP_LIBS="libdsrt_unord.so,libdsrt_ns3u.so,libdsrt_ns4u.so,libdsrt_cs3u.so,libdsrt_unord.so,libdsrt_ns3u.so,libdsrt_ns4u.so,libdsrt_cs3u.so" SWEEP_OPTS='[{}]'
SWEEP_SCENE=c2 python tools/sweeps/sweep_variants.py 64 > $O/r2c38_c2.log 2>&1; cat $O/r2c38_c2.log
export SWEEP_LIBS="libdsrt_unord.so,libdsrt_ns3u.so,libdsrt_ns4u.so,libdsrt_cs3u.so"
SWEEP_SCENE=c3 python tools/sweeps/sweep_variants.py 64 > $O/r2c38_c3.log 2>&1; cat $O/r2c38_c3.log
SWEEP_SCENE=soup1 python tools/sweeps/sweep_variants.py 8 > $O/r2c38_soup1.log 2>&1; cat $O/r2c38_soup1.log
