#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
L="libdsrt.so,libdsrt_p128.so,libdsrt_p96.so,libdsrt_s0.so,libdsrt_oh.so,libdsrt.so"
SWEEP_SCENE=soup8 SWEEP_LIBS=$L SWEEP_OPTS='[{}]' python tools/sweeps/sweep_variants.py 8 > $O/r2c10_sweep_soup8.log 2>&1; cat $O/r2c10_sweep_soup8.log
SWEEP_SCENE=c2 SWEEP_LIBS=$L SWEEP_OPTS='[{}]' python tools/sweeps/sweep_variants.py 64 > $O/r2c10_sweep_c2.log 2>&1; cat $O/r2c10_sweep_c2.log
SWEEP_SCENE=soup64 SWEEP_DEVICE_BUILD=1 SWEEP_LIBS="libdsrt.so,libdsrt_p96.so" SWEEP_OPTS='[{}]' python tools/sweeps/sweep_variants.py 8 > $O/r2c10_sweep_soup64.log 2>&1; cat $O/r2c10_sweep_soup64.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
