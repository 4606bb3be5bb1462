#!/usr/bin/env bash
# the node's flat word taken from the node's first 256-bit load (NodeRegs::flat) instead of a separate load in drop_source's branch
set -u
O=gpurun_out
export SWEEP_LIBS="libdsrt_prev.so,libdsrt.so,libdsrt_prev.so,libdsrt.so" SWEEP_OPTS='[{}]'
SWEEP_SCENE=c2 python tools/sweeps/sweep_variants.py 64 > $O/r2c34_c2.log 2>&1; cat $O/r2c34_c2.log
SWEEP_SCENE=c3 python tools/sweeps/sweep_variants.py 64 > $O/r2c34_c3.log 2>&1; cat $O/r2c34_c3.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c34_pytest.log 2>&1; tail -2 $O/r2c34_pytest.log
