#!/usr/bin/env bash
# tools/measure_r2.sh without the 64 Mi host build and the builder comparison (those take 4 of its 7.5 minutes):
# bash tools/measure_r2_short.sh [tag]
set -u
O=gpurun_out; T=${1:-r2}; mkdir -p $O
export DSRT_PARITY_LOG=$PWD/$O/parity_$T.jsonl; rm -f $DSRT_PARITY_LOG
python -m pytest tests -m gpu -q 2>&1 | tail -3
unset DSRT_PARITY_LOG
python bench.py --steps 3 --warmup 3 > $O/bench_$T.json 2> $O/bench_$T.err; cut -c1-220 $O/bench_$T.json
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --skip-null-shadow > $O/bench_skipnull_$T.json 2> $O/bench_skipnull_$T.err; cut -c1-160 $O/bench_skipnull_$T.json
python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref_$T.json 2> $O/bench_ref_$T.err; cut -c1-200 $O/bench_ref_$T.json
python bench.py --spp 16 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$T.csv python bench.py --spp 16 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_list_$T.log 2>&1
python profiles/profile_run.py 4 > $O/pr_on_$T.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_trace -s 18 -c 2 -f -o $O/prof_on_$T python profiles/profile_run.py 4 > $O/pr_ncu_on_$T.log 2>&1
OFF="refill_busy_lanes=0 refill_hi_lanes=0 postpone_min_lanes=0 coop_min_pairs=1000000"
python profiles/profile_run.py 4 $OFF > $O/pr_off_$T.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_trace -s 18 -c 2 -f -o $O/prof_off_$T python profiles/profile_run.py 4 $OFF > $O/pr_ncu_off_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_shade|k_generate" -s 10 -c 2 -f -o $O/prof_shade_$T python profiles/profile_run.py 4 > $O/pr_ncu_shade_$T.log 2>&1
cat $O/pr_on_$T.log $O/pr_off_$T.log
rm -f $O/cfg_$T.jsonl
for w in c3 c4 soup1 soup8; do
  python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline >> $O/cfg_$T.jsonl 2>> $O/cfg_$T.err; tail -1 $O/cfg_$T.jsonl | cut -c1-160
done
for w in soup1 soup8 soup64; do
  python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --device-build >> $O/cfg_$T.jsonl 2>> $O/cfg_$T.err; tail -1 $O/cfg_$T.jsonl | cut -c1-160
done
cp dsgpuraytracing_b200/csrc/build.log $O/build_$T.log
