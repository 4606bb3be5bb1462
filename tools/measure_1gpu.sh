#!/usr/bin/env bash
# The round's 1-GPU measurement pass (run on a B200 box from the repo root; everything lands in gpurun_out/):
# GPU tests, bench line, ncu launch list of the bench command, ncu --set full of the depth-0 traversal launches with and
# without compaction, k_shade / k_generate capture, the BASELINE configurations, CPU reference modes, reference arm.
set -u
O=gpurun_out; T=${1:-final}
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 > $O/bench_$T.json 2> $O/bench_$T.err; cut -c1-220 $O/bench_$T.json
python bench.py --spp 16 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$T.csv python bench.py --spp 16 --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_list_$T.log 2>&1
python profiles/profile_run.py 4 > $O/pr_on_$T.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_trace -s 18 -c 2 -o $O/prof_on_$T python profiles/profile_run.py 4 > $O/pr_ncu_on_$T.log 2>&1
OFF="refill_busy_lanes=0 postpone_min_lanes=0 coop_min_pairs=1000000"
python profiles/profile_run.py 4 $OFF > $O/pr_off_$T.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_trace -s 18 -c 2 -o $O/prof_off_$T python profiles/profile_run.py 4 $OFF > $O/pr_ncu_off_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_shade|k_generate" -s 10 -c 2 -o $O/prof_shade_$T python profiles/profile_run.py 4 > $O/pr_ncu_shade_$T.log 2>&1
cat $O/pr_on_$T.log $O/pr_off_$T.log
python tools/run_configs.py --soup-max 64 > $O/cfg1_$T.jsonl 2> $O/cfg1_$T.err; wc -l $O/cfg1_$T.jsonl
python tools/cpu_reference_modes.py > $O/cpu_modes_$T.json 2> $O/cpu_modes_$T.err
python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref_$T.json 2> $O/bench_ref_$T.err; cut -c1-200 $O/bench_ref_$T.json
