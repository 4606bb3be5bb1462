#!/usr/bin/env bash
# coplanar slot mates dropped with the source + light-aligned grid (+ regroup_top as an option): GPU tests, A/B in one process (c2, c3), bench line
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c32_pytest.log 2>&1; tail -2 $O/r2c32_pytest.log
timeout 600 python tools/sweeps/sweep_light_grid.py 64 > $O/r2c32_tree.log 2>&1; cat $O/r2c32_tree.log
timeout 600 python bench.py > $O/r2c32_bench.json 2> $O/r2c32_bench.err; cat $O/r2c32_bench.json | cut -c1-300
