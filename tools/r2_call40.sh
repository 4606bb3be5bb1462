#!/usr/bin/env bash
# new defaults (unordered any-hit children, 4 node steps): GPU tests; 7 vs 8 resident CTAs per SM (72 / 64 registers, no spills in either)
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c40_pytest.log 2>&1; tail -3 $O/r2c40_pytest.log
export SWEEP_LIBS="libdsrt.so,libdsrt_c8.so,libdsrt.so,libdsrt_c8.so" SWEEP_OPTS='[{}]'
SWEEP_SCENE=c2 python tools/sweeps/sweep_variants.py 64 > $O/r2c40_c2.log 2>&1; cat $O/r2c40_c2.log
export SWEEP_LIBS="libdsrt.so,libdsrt_c8.so"
SWEEP_SCENE=c3 python tools/sweeps/sweep_variants.py 64 > $O/r2c40_c3.log 2>&1; cat $O/r2c40_c3.log
SWEEP_SCENE=soup8 python tools/sweeps/sweep_variants.py 8 > $O/r2c40_soup8.log 2>&1; cat $O/r2c40_soup8.log
