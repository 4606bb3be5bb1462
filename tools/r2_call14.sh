#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2c14_bench.json 2> $O/r2c14_bench.err; cut -c1-200 $O/r2c14_bench.json
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --skip-null-shadow > $O/r2c14_bench_skipnull.json 2> $O/r2c14_bench_skipnull.err; cut -c1-200 $O/r2c14_bench_skipnull.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload c3 > $O/r2c14_bench_c3.json 2>/dev/null; cut -c1-200 $O/r2c14_bench_c3.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload c3 --skip-null-shadow > $O/r2c14_bench_c3_skipnull.json 2>/dev/null; cut -c1-200 $O/r2c14_bench_c3_skipnull.json
