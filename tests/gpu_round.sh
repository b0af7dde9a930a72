#!/bin/bash
O=gpurun_out
echo "== filter case plain"; timeout 100 python tests/filter_case.py a.0+3.4 2>&1 | tail -3
echo "== filter case under memcheck"; timeout 400 compute-sanitizer --tool memcheck --print-limit 3 python tests/filter_case.py a.0+3.4 > $O/memcheck.log 2>&1; grep -v "^=========     Host Frame\|^=========         " $O/memcheck.log | head -40 | cut -c1-250
echo "== ncu decode (current decoder)"; AB_TAG=ncu timeout 300 ncu --clock-control none --import-source on --section SourceCounters --section WarpStateStats --section SchedulerStats -k regex:xzb_k_decode -c 1 -f -o $O/r02_decode2 python tests/ab_probe.py T,6,4194304,4194304 > $O/ncu_dec.log 2>&1; tail -2 $O/ncu_dec.log
