#!/bin/bash
O=gpurun_out
echo "== probe quick"; timeout 200 python tests/gpu_probe_dp.py quick > $O/probe_q.log 2>&1; echo rc=$? >> $O/probe_q.log; grep -c "^OK" $O/probe_q.log; grep -v "^OK" $O/probe_q.log | tail -4
echo "== pytest subset (decoder verdicts, fast presets, filters)"; timeout 420 python -m pytest tests/test_gpu_parity.py tests/test_gpu_filters.py -q -k "small or corpus or truncated or fast_presets or chain or bad_chains" > $O/pytest_subset.log 2>&1; tail -5 $O/pytest_subset.log | cut -c1-300
echo "== filter case under memcheck (short)"; timeout 300 compute-sanitizer --tool memcheck --print-limit 3 python tests/filter_case.py a.0+3.4 300000 > $O/memcheck.log 2>&1; grep -v "Host Frame\|^=========         " $O/memcheck.log | head -12 | cut -c1-200
echo "== prof"; XZB_OVERLAP=0 timeout 120 python tests/prof_dp.py > $O/prof_a.log 2>&1; grep "DPPROF chain\|DPPROF nodes\|DPPROF loop\|OK\|MISM" $O/prof_a.log | cut -c1-330
echo "== timing"; AB_TAG=new timeout 200 python tests/ab_probe.py T,6,33554432,4194304 E,6,33554432,4194304 2>&1 | tail -3
