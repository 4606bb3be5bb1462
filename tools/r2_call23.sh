#!/usr/bin/env bash
# 96-byte nodes fetched with three 256-bit loads (+ float scales) vs 80-byte nodes / five 128-bit loads
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c23_pytest.log 2>&1; tail -2 $O/r2c23_pytest.log
export SWEEP_LIBS="libdsrt_n1.so,libdsrt_n80.so,libdsrt.so,libdsrt_n80.so,libdsrt.so" SWEEP_OPTS='[{}]'
python tools/sweeps/sweep_variants.py 64 > $O/r2c23_sweep_c2.log 2>&1; cat $O/r2c23_sweep_c2.log
export SWEEP_LIBS="libdsrt_n80.so,libdsrt.so"
SWEEP_SCENE=soup8 python tools/sweeps/sweep_variants.py 16 > $O/r2c23_sweep_soup8.log 2>&1; cat $O/r2c23_sweep_soup8.log
SWEEP_SCENE=soup1 python tools/sweeps/sweep_variants.py 16 > $O/r2c23_sweep_soup1.log 2>&1; cat $O/r2c23_sweep_soup1.log
