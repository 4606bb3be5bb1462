#!/usr/bin/env bash
set -u
O=gpurun_out; mkdir -p $O
python tools/sweeps/sweep_collapse_soup.py 8 8 > $O/r2c13_collapse_soup8.log 2>&1; cat $O/r2c13_collapse_soup8.log
python tools/sweeps/sweep_collapse_soup.py 1 8 > $O/r2c13_collapse_soup1.log 2>&1; cat $O/r2c13_collapse_soup1.log
