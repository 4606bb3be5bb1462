#!/usr/bin/env bash
# run-time knobs on the final build: batch size, pool group, refill / patience, node steps (compiled variants ns1 / ns3)
set -u
O=gpurun_out
export SWEEP_LIBS="libdsrt.so" SWEEP_OPTS='[{}, {"batch_spp": 4}, {"batch_spp": 16}, {"pool_batches": 4}, {"pool_batches": 16}, {"refill_hi_lanes": 22}, {"refill_hi_lanes": 28}, {"refill_patience": 4}, {"refill_patience": 10}, {"postpone_wait_mode": 1}, {"postpone_wait_mode": 4}]'
python tools/sweeps/sweep_variants.py 64 > $O/r2c28_sweep_knobs.log 2>&1; cat $O/r2c28_sweep_knobs.log
export SWEEP_LIBS="libdsrt.so,libdsrt_ns1.so,libdsrt_ns3.so,libdsrt_pf0.so,libdsrt_pf64k.so" SWEEP_OPTS='[{}]'
python tools/sweeps/sweep_variants.py 64 > $O/r2c28_sweep_builds.log 2>&1; cat $O/r2c28_sweep_builds.log
