#!/usr/bin/env bash
# Copies the outputs of tools/measure_1gpu.sh <tag> from gpurun_out/ into profiles/ (tracked) and regenerates the derived
# tables: bash tools/collect_profiles.sh <tag>
set -eu
T=${1:-final}; O=gpurun_out; P=profiles
ncu -i $O/prof_on_$T.ncu-rep --page raw --csv > $P/r1_ktrace_depth0_raw.csv 2>/dev/null
ncu -i $O/prof_off_$T.ncu-rep --page raw --csv > $P/r1_ktrace_depth0_nocompaction_raw.csv 2>/dev/null
ncu -i $O/prof_shade_$T.ncu-rep --page raw --csv > $P/r1_kshade_raw.csv 2>/dev/null
cp $O/launches_$T.csv $P/r1_launches_bench_spp16.csv
cp $O/bench_$T.json $P/r1_bench_1gpu.json
cp $O/bench_ref_$T.json $P/r1_bench_reference_arm.json
cp $O/cpu_modes_$T.json $P/r1_cpu_reference_modes.json
cp $O/cfg1_$T.jsonl $P/r1_configs.jsonl
python tools/profile_tables.py $P/r1_ktrace_depth0_raw.csv $P/r1_ktrace_depth0_nocompaction_raw.csv $P/r1_bench_1gpu.json > $P/r1_compaction_table.md
(echo "# python tools/launch_shares.py profiles/r1_launches_bench_spp16.csv"; python tools/launch_shares.py $P/r1_launches_bench_spp16.csv) > $P/r1_launch_shares.txt
cat $P/r1_compaction_table.md; head -8 $P/r1_launch_shares.txt
