#!/usr/bin/env bash
# primitive bits enumerated lowest first (one slow-pipe op per pair in the scatter loop) vs the measured build; optional cap on an owner's pairs per round
set -u
O=gpurun_out
export SWEEP_LIBS="libdsrt_prev.so,libdsrt.so,libdsrt_cap3.so,libdsrt_cap4.so,libdsrt_prev.so,libdsrt.so" SWEEP_OPTS='[{}]'
python tools/sweeps/sweep_variants.py 64 > $O/r2c25_sweep_c2.log 2>&1; cat $O/r2c25_sweep_c2.log
export SWEEP_LIBS="libdsrt_prev.so,libdsrt.so,libdsrt_cap3.so"
SWEEP_SCENE=soup8 python tools/sweeps/sweep_variants.py 16 > $O/r2c25_sweep_soup8.log 2>&1; cat $O/r2c25_sweep_soup8.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c25_pytest.log 2>&1; tail -2 $O/r2c25_pytest.log
