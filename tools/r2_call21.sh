#!/usr/bin/env bash
# nibble-per-slot hit mask (new node meta words) vs the previous build; e2e host trace with the cudaMemGetInfo fast path
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c21_pytest.log 2>&1; tail -3 $O/r2c21_pytest.log
export SWEEP_LIBS="libdsrt_old.so,libdsrt.so,libdsrt_unord.so,libdsrt_old.so,libdsrt.so" SWEEP_OPTS='[{}]'
python tools/sweeps/sweep_variants.py 64 > $O/r2c21_sweep_c2.log 2>&1; cat $O/r2c21_sweep_c2.log
export SWEEP_LIBS="libdsrt_old.so,libdsrt.so,libdsrt_unord.so"
SWEEP_SCENE=soup8 python tools/sweeps/sweep_variants.py 16 > $O/r2c21_sweep_soup8.log 2>&1; cat $O/r2c21_sweep_soup8.log
DSRT_HOST_TRACE=1 python tools/e2e_jitter.py 14 > $O/r2c21_e2e_trace.log 2>&1; grep -c "host trace" $O/r2c21_e2e_trace.log; grep "^step" $O/r2c21_e2e_trace.log
