#!/usr/bin/env bash
# device-side builder with the light-aligned grid and k_db_mark_flat: GPU tests + host vs device builder on the bench scene / 1 Mi soup
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c35_pytest.log 2>&1; tail -3 $O/r2c35_pytest.log
DSRT_BUILD_TIMING=1 timeout 600 python tools/device_build_bench.py 0 1 > $O/r2c35_devbuild.jsonl 2> $O/r2c35_devbuild.err; cut -c1-330 $O/r2c35_devbuild.jsonl
