#!/usr/bin/env bash
# occupancy: 8 CTAs/SM at 64 registers (no spills since the node test shrank) and 9 at 56, with and without the smaller shared-memory footprint
set -u
O=gpurun_out
export SWEEP_LIBS="libdsrt.so,libdsrt_c8.so,libdsrt_c8s.so,libdsrt_c9s.so,libdsrt.so" SWEEP_OPTS='[{}]'
python tools/sweeps/sweep_variants.py 64 > $O/r2c24_sweep_c2.log 2>&1; cat $O/r2c24_sweep_c2.log
export SWEEP_LIBS="libdsrt.so,libdsrt_c8.so,libdsrt_c8s.so,libdsrt_c9s.so"
SWEEP_SCENE=soup8 python tools/sweeps/sweep_variants.py 16 > $O/r2c24_sweep_soup8.log 2>&1; cat $O/r2c24_sweep_soup8.log
