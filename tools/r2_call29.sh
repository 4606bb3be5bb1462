#!/usr/bin/env bash
# surplus warps of k_trace leave before the work counter on short queues: per-launch durations of one 8-spp batch + its pool iterations (ncu list, cold) and whole frames
set -u
O=gpurun_out
for L in libdsrt_prev.so libdsrt.so; do
  DSRT_LIB=$PWD/dsgpuraytracing_b200/$L ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r2c29_launches_$L.csv python profiles/profile_run.py 8 > /dev/null 2>&1
done
export SWEEP_LIBS="libdsrt_prev.so,libdsrt.so,libdsrt_prev.so,libdsrt.so" SWEEP_OPTS='[{}, {"batch_spp": 4, "pool_batches": 1}]'
python tools/sweeps/sweep_variants.py 64 > $O/r2c29_sweep_c2.log 2>&1; cat $O/r2c29_sweep_c2.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c29_pytest.log 2>&1; tail -2 $O/r2c29_pytest.log
