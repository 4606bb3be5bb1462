#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "device_built" 2>&1 | tail -25 > $O/r2c7_pytest.log; tail -25 $O/r2c7_pytest.log
DSRT_BUILD_TIMING=1 timeout 900 python tools/device_build_bench.py 0 1 8 > $O/r2c7_devbuild.jsonl 2> $O/r2c7_devbuild.err; cat $O/r2c7_devbuild.jsonl; grep -E "device_build|dsrt_build_bvh2|build_wide" $O/r2c7_devbuild.err | tail -12
