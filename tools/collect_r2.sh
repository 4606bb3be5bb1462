#!/usr/bin/env bash
# Copies the outputs of tools/measure_r2.sh <tag> from gpurun_out/ into profiles/ (tracked) and regenerates the derived
# tables: bash tools/collect_r2.sh <tag>
set -eu
T=${1:-r2}; O=gpurun_out; P=profiles
ncu -i $O/prof_on_$T.ncu-rep --page raw --csv > $P/r2_ktrace_depth0_raw.csv 2>/dev/null
ncu -i $O/prof_off_$T.ncu-rep --page raw --csv > $P/r2_ktrace_depth0_nocompaction_raw.csv 2>/dev/null
ncu -i $O/prof_shade_$T.ncu-rep --page raw --csv > $P/r2_kshade_raw.csv 2>/dev/null
cp $O/launches_$T.csv $P/r2_launches_bench_spp16.csv
cp $O/bench_$T.json $P/r2_bench_1gpu.json
cp $O/bench_skipnull_$T.json $P/r2_bench_1gpu_skip_null_shadow.json
cp $O/bench_ref_$T.json $P/r2_bench_reference_arm.json
cp $O/cfg_$T.jsonl $P/r2_configs.jsonl
cp $O/devbuild_$T.jsonl $P/r2_device_build.jsonl
grep -E "device_build:|dsrt_build_bvh2:|build_wide_bvh:" $O/devbuild_$T.err > $P/r2_device_build_stages.txt || true
cp $O/parity_$T.jsonl $P/r2_parity.jsonl
python tools/parity_table.py $P/r2_parity.jsonl > $P/r2_parity.md
python tools/profile_tables.py $P/r2_ktrace_depth0_raw.csv $P/r2_ktrace_depth0_nocompaction_raw.csv $P/r2_bench_1gpu.json > $P/r2_compaction_table.md
(echo "# python tools/launch_shares.py profiles/r2_launches_bench_spp16.csv"; python tools/launch_shares.py $P/r2_launches_bench_spp16.csv) > $P/r2_launch_shares.txt
# SASS of both k_trace instantiations (the build that was measured) + per-source-line issue breakdown of the connect kernel
ncu -i $O/prof_on_$T.ncu-rep --page source --csv --print-source sass > /tmp/r2_src.csv 2>/dev/null
rm -rf /tmp/xelf_r2 && mkdir /tmp/xelf_r2 && (cd /tmp/xelf_r2 && cuobjdump -xelf all $OLDPWD/dsgpuraytracing_b200/libdsrt.so > /dev/null && nvdisasm -g -c dsrt_api.sm_100a.cubin > /tmp/r2_dis.txt 2>/dev/null)
# (cuobjdump prints every function; keep the two production instantiations, without the encoding column)
cuobjdump -sass dsgpuraytracing_b200/libdsrt.so | sed -E 's/\s+\/\* 0x[0-9a-f]{16} \*\///' | awk '/Function :/ {keep = ($0 ~ /k_traceILb1ELb0/)} keep && (/Function :/ || /^[ \t]+\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f]\*\//)' > $P/r2_ktrace_any_sass.txt
cuobjdump -sass dsgpuraytracing_b200/libdsrt.so | sed -E 's/\s+\/\* 0x[0-9a-f]{16} \*\///' | awk '/Function :/ {keep = ($0 ~ /k_traceILb0ELb0/)} keep && (/Function :/ || /^[ \t]+\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f]\*\//)' > $P/r2_ktrace_closest_sass.txt
python tools/sass_by_line.py /tmp/r2_src.csv /tmp/r2_dis.txt 'k_traceILb1ELb0' 2 60 > $P/r2_ktrace_connect_by_line.txt || true
cat $P/r2_compaction_table.md; head -8 $P/r2_launch_shares.txt; head -12 $P/r2_ktrace_connect_by_line.txt
# the same counters grouped by what the code does (both instantiations)
(python tools/sass_categories.py /tmp/r2_src.csv /tmp/r2_dis.txt k_traceILb1ELb0 2; python tools/sass_categories.py /tmp/r2_src.csv /tmp/r2_dis.txt k_traceILb0ELb0 0) > $P/r2_ktrace_categories_raw.txt || true
