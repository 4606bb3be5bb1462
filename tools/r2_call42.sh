#!/usr/bin/env bash
# final build (3 node steps in the closest-hit kernel): GPU tests, smoke, bench line
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c42_pytest.log 2>&1; tail -2 $O/r2c42_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/r2c42_smoke.log 2>&1; tail -2 $O/r2c42_smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > $O/r2c42_bench.json 2> $O/r2c42_bench.err; cut -c1-250 $O/r2c42_bench.json
