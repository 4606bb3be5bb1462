#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
SWEEP_LIBS="libdsrt.so,libdsrt_ns2.so,libdsrt_ns3.so,libdsrt_ns2p96.so,libdsrt_ns2s0.so" SWEEP_OPTS='[{}, {"postpone_min_lanes": 8}, {"postpone_min_lanes": 16}, {"refill_busy_lanes": 14}, {"refill_busy_lanes": 22}, {"coop_min_pairs": 12}, {"postpone_wait_mode": 1}]' \
  python tools/sweeps/sweep_variants.py 64 > $O/r2c5_sweep.log 2>&1
cat $O/r2c5_sweep.log
timeout 600 python -m pytest tests -m gpu -q -k "shim or skip_null or mirror or many_light" 2>&1 | tail -15 > $O/r2c5_pytest.log; cat $O/r2c5_pytest.log
DSRT_LIB=$PWD/dsgpuraytracing_b200/libdsrt_ns2.so timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $O/r2c5_pytest_ns2.log; cat $O/r2c5_pytest_ns2.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-null-shadow > $O/r2c5_bench_skipnull.json 2> $O/r2c5_bench_skipnull.err; cut -c1-400 $O/r2c5_bench_skipnull.json
