#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
L="libdsrt_nosatc.so,libdsrt.so,libdsrt_nosatc.so,libdsrt.so"
SWEEP_SCENE=c2 SWEEP_LIBS=$L SWEEP_OPTS='[{}]' python tools/sweeps/sweep_variants.py 64 > $O/r2c15_sweep_c2.log 2>&1; cat $O/r2c15_sweep_c2.log
SWEEP_SCENE=soup8 SWEEP_LIBS=$L SWEEP_OPTS='[{}]' python tools/sweeps/sweep_variants.py 8 > $O/r2c15_sweep_soup8.log 2>&1; cat $O/r2c15_sweep_soup8.log
export DSRT_PARITY_LOG=$PWD/$O/r2c15_parity.jsonl; rm -f $DSRT_PARITY_LOG
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
grep '"gate": "ids"' $DSRT_PARITY_LOG | cut -c1-200
