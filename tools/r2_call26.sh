#!/usr/bin/env bash
# repeatability of the lowest-bit-first build against the previous one (alternating, three runs each, then 256 spp once each)
set -u
O=gpurun_out
export SWEEP_LIBS="libdsrt.so,libdsrt_prev.so,libdsrt.so,libdsrt_prev.so,libdsrt.so,libdsrt_prev.so" SWEEP_OPTS='[{}, {}]'
python tools/sweeps/sweep_variants.py 64 > $O/r2c26_sweep_c2.log 2>&1; cat $O/r2c26_sweep_c2.log
export SWEEP_LIBS="libdsrt.so,libdsrt_prev.so" SWEEP_OPTS='[{}]'
python tools/sweeps/sweep_variants.py 256 > $O/r2c26_sweep_c2_256.log 2>&1; cat $O/r2c26_sweep_c2_256.log
