#!/usr/bin/env bash
# round 2, GPU call 1: A/B of the node-test / triangle-test variants + GPU tests + bench line
set -u
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $O/r2c1_gpu.txt
SWEEP_LIBS="libdsrt_base.so,libdsrt.so,libdsrt_sat.so,libdsrt_fast.so,libdsrt_ns2.so,libdsrt_c8.so,libdsrt_c6.so,libdsrt_base.so,libdsrt.so" SWEEP_OPTS='[{}]' \
  python tools/sweeps/sweep_variants.py 64 > $O/r2c1_sweep.log 2>&1
cat $O/r2c1_sweep.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/r2c1_pytest.log; cat $O/r2c1_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2c1_bench.json 2> $O/r2c1_bench.err; cut -c1-300 $O/r2c1_bench.json
