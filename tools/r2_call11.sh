#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
L="libdsrt.so,libdsrt_p128s0.so,libdsrt_p96s0.so,libdsrt_p64s0.so,libdsrt.so"
SWEEP_SCENE=c2 SWEEP_LIBS=$L SWEEP_OPTS='[{}, {"smem_carveout_pct": 58}, {"smem_carveout_pct": 72}]' python tools/sweeps/sweep_variants.py 64 > $O/r2c11_sweep_c2.log 2>&1; cat $O/r2c11_sweep_c2.log
SWEEP_SCENE=soup8 SWEEP_LIBS=$L SWEEP_OPTS='[{}]' python tools/sweeps/sweep_variants.py 8 > $O/r2c11_sweep_soup8.log 2>&1; cat $O/r2c11_sweep_soup8.log
