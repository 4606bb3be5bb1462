#!/usr/bin/env bash
set -u
O=gpurun_out; T=r2
OFF="refill_busy_lanes=0 refill_hi_lanes=0 postpone_min_lanes=0 coop_min_pairs=1000000"
python profiles/profile_run.py 4 $OFF > $O/pr_off_$T.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_trace -s 18 -c 2 -f -o $O/prof_off_$T python profiles/profile_run.py 4 $OFF > $O/pr_ncu_off_$T.log 2>&1
cat $O/pr_off_$T.log
