#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/r2c2_pytest.log; cat $O/r2c2_pytest.log
python profiles/profile_run.py 4 > $O/r2c2_pr.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_trace -s 18 -c 2 -f -o $O/r2c2_prof python profiles/profile_run.py 4 > $O/r2c2_ncu.log 2>&1
cat $O/r2c2_pr.log; tail -3 $O/r2c2_ncu.log; ls -la $O/r2c2_prof.ncu-rep
