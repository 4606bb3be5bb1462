#!/usr/bin/env bash
# light-aligned quantisation grid: GPU tests, A/B in one process (c2, c3), bench line
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c30_pytest.log 2>&1; tail -2 $O/r2c30_pytest.log
timeout 600 python tools/sweeps/sweep_light_grid.py 64 > $O/r2c30_light_grid.log 2>&1; cat $O/r2c30_light_grid.log
timeout 600 python bench.py > $O/r2c30_bench.json 2> $O/r2c30_bench.err; cat $O/r2c30_bench.json | cut -c1-400
