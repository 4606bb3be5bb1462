#!/usr/bin/env bash
# compile-time variants re-checked on the refined tree: 3 node steps per primitive decision, unordered any-hit children
set -u
O=gpurun_out
export SWEEP_LIBS="libdsrt.so,libdsrt_ns3.so,libdsrt_unord.so,libdsrt.so,libdsrt_ns3.so,libdsrt_unord.so" SWEEP_OPTS='[{}]'
SWEEP_SCENE=c2 python tools/sweeps/sweep_variants.py 64 > $O/r2c37_c2.log 2>&1; cat $O/r2c37_c2.log
SWEEP_SCENE=c3 python tools/sweeps/sweep_variants.py 64 > $O/r2c37_c3.log 2>&1; cat $O/r2c37_c3.log
