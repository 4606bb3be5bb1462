#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
DSRT_BUILD_TIMING=1 python tools/device_build_bench.py 8 64 > $O/r2c12_devbuild.jsonl 2> $O/r2c12_devbuild.err; cut -c1-300 $O/r2c12_devbuild.jsonl; grep -E "device_build:|dsrt_build_accel:" $O/r2c12_devbuild.err
