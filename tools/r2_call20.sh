#!/usr/bin/env bash
# e2e host-phase trace + two-streams-on-one-GPU probe
set -u
O=gpurun_out
DSRT_HOST_TRACE=1 python tools/e2e_jitter.py 14 > $O/e2e_trace.log 2>&1
python tools/sweeps/overlap_probe.py 256 > $O/overlap_probe.log 2>&1
DSRT_STAGGER=1 python tools/sweeps/overlap_probe.py 256 > $O/overlap_probe_stagger.log 2>&1
tail -5 $O/e2e_trace.log; cat $O/overlap_probe.log $O/overlap_probe_stagger.log
