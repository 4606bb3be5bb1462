#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
SWEEP_LIBS="libdsrt_r1.so,libdsrt_nopf.so,libdsrt_nooh.so,libdsrt.so,libdsrt_pf4k.so,libdsrt_pf64k.so,libdsrt_ns2.so,libdsrt_c8.so,libdsrt_r1.so,libdsrt.so" SWEEP_OPTS='[{}]' \
  python tools/sweeps/sweep_variants.py 64 > $O/r2c4_sweep.log 2>&1
cat $O/r2c4_sweep.log
export DSRT_PARITY_LOG=$PWD/$O/r2c4_parity.jsonl; rm -f $DSRT_PARITY_LOG
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $O/r2c4_pytest.log; tail -12 $O/r2c4_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2c4_bench.json 2> $O/r2c4_bench.err; cut -c1-250 $O/r2c4_bench.json; tail -3 $O/r2c4_bench.err
