#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
SWEEP_SCENE=soup8 SWEEP_LIBS="libdsrt.so,libdsrt_r1like.so,libdsrt_nooh.so,libdsrt_nofast.so,libdsrt_ns1.so,libdsrt_p96.so,libdsrt_nosat.so,libdsrt_nopf.so" SWEEP_OPTS='[{}, {"postpone_min_lanes": 12}]' \
  python tools/sweeps/sweep_variants.py 8 > $O/r2c9_sweep_soup8.log 2>&1
cat $O/r2c9_sweep_soup8.log
timeout 600 python -m pytest tests -m gpu -q -k "window or invalidate" 2>&1 | tail -5
