#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/r2c6_pytest.log; tail -12 $O/r2c6_pytest.log
python bench.py --steps 3 --warmup 3 > $O/r2c6_bench.json 2> $O/r2c6_bench.err; cut -c1-250 $O/r2c6_bench.json; tail -3 $O/r2c6_bench.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
