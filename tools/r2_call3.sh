#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
export DSRT_PARITY_LOG=$PWD/$O/r2c3_parity.jsonl; rm -f $DSRT_PARITY_LOG
timeout 1200 python -m pytest tests -m gpu -q -k "not cblucy" 2>&1 | tail -8 > $O/r2c3_pytest.log; cat $O/r2c3_pytest.log
python profiles/profile_run.py 4 > $O/r2c3_pr.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_trace -s 18 -c 2 -f -o $O/r2c3_prof python profiles/profile_run.py 4 > $O/r2c3_ncu.log 2>&1
cat $O/r2c3_pr.log
python bench.py --steps 3 --warmup 3 > $O/r2c3_bench.json 2> $O/r2c3_bench.err; cut -c1-250 $O/r2c3_bench.json; tail -3 $O/r2c3_bench.err
