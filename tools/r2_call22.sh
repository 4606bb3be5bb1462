#!/usr/bin/env bash
# micro batch (pre-biased exponents, oct4, cheap source check, no redundant source compare) vs the nibble build; ncu capture
set -u
O=gpurun_out; T=r2n
export SWEEP_LIBS="libdsrt_n1.so,libdsrt.so,libdsrt_n1.so,libdsrt.so" SWEEP_OPTS='[{}]'
python tools/sweeps/sweep_variants.py 64 > $O/r2c22_sweep_c2.log 2>&1; cat $O/r2c22_sweep_c2.log
export SWEEP_LIBS="libdsrt_n1.so,libdsrt.so"
SWEEP_SCENE=soup8 python tools/sweeps/sweep_variants.py 16 > $O/r2c22_sweep_soup8.log 2>&1; cat $O/r2c22_sweep_soup8.log
export SWEEP_LIBS="libdsrt.so" SWEEP_OPTS='[{}, {"postpone_min_lanes": 6}, {"postpone_min_lanes": 10}, {"refill_busy_lanes": 16}, {"refill_busy_lanes": 20}, {"coop_min_pairs": 4}, {"coop_min_pairs": 9}]'
python tools/sweeps/sweep_variants.py 64 > $O/r2c22_sweep_knobs.log 2>&1; cat $O/r2c22_sweep_knobs.log
python profiles/profile_run.py 4 > $O/pr_on_$T.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_trace -s 18 -c 2 -f -o $O/prof_on_$T python profiles/profile_run.py 4 > $O/pr_ncu_on_$T.log 2>&1
cat $O/pr_on_$T.log
