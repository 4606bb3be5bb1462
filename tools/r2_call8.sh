#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "device_built" 2>&1 | tail -5 > $O/r2c8_pytest.log; tail -5 $O/r2c8_pytest.log
for lib in libdsrt_a10.so libdsrt.so libdsrt_a16.so; do
  echo "== $lib"; DSRT_LIB=$PWD/dsgpuraytracing_b200/$lib DSRT_BUILD_TIMING=1 timeout 900 python tools/device_build_bench.py 0 1 8 2> $O/r2c8_$lib.err | grep device_lbvh | cut -c1-330
done
