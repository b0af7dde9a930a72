#!/bin/bash
# ncu source-level capture of the DP parser kernel on one 1 MiB T block (run under gpurun)
O=gpurun_out
ncu --clock-control none --import-source on --section SourceCounters --section WarpStateStats --section SchedulerStats --section SpeedOfLight --section LaunchStats --section Occupancy \
  -k regex:xzb_k_parse_dp -c 1 -f -o $O/${1:-r02_dp} python tests/ab_probe.py T,6,1048576,1048576 > $O/ncu_dp.log 2>&1
tail -3 $O/ncu_dp.log
