#!/bin/bash
# Development round on the GPU box (run under gpurun): every step under its own timeout.
O=gpurun_out
echo "== prof lag1"; XZB_OVERLAP=0 timeout 120 python tests/prof_dp.py > $O/prof_a.log 2>&1; grep "DPPROF chain\|DPPROF nodes\|OK\|MISM" $O/prof_a.log | cut -c1-330
echo "== prof lag2"; XZB200_LIB=xz_b200/libxzb200_prof_lag2.so XZB_OVERLAP=0 timeout 120 python tests/prof_dp.py > $O/prof_b.log 2>&1; grep "DPPROF chain\|DPPROF nodes\|OK\|MISM" $O/prof_b.log | cut -c1-330
echo "== ncu decode"; timeout 300 ncu --clock-control none --import-source on --section SourceCounters --section WarpStateStats --section SchedulerStats --section SpeedOfLight --section LaunchStats --section Occupancy -k regex:xzb_k_decode -c 1 -f -o $O/r02_decode python tests/ab_probe.py T,6,4194304,4194304 > $O/ncu_dec.log 2>&1; tail -2 $O/ncu_dec.log
