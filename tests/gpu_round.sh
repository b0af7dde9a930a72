#!/bin/bash
# Development round on the GPU box (run under gpurun): correctness probe, DP profile A/B, decoder A/B.
O=gpurun_out
python tests/gpu_probe_dp.py quick > $O/probe_q.log 2>&1; echo rc=$? >> $O/probe_q.log
XZB_OVERLAP=0 python tests/prof_dp.py > $O/prof_a.log 2>&1
XZB200_LIB=xz_b200/libxzb200_prof_pf.so XZB_OVERLAP=0 python tests/prof_dp.py > $O/prof_b.log 2>&1
XZB200_LIB=xz_b200/libxzb200_prof_lag2.so XZB_OVERLAP=0 python tests/prof_dp.py > $O/prof_c.log 2>&1
AB_TAG=new python tests/ab_probe.py T,6,33554432,4194304 E,6,33554432,4194304 R,3,33554432,4194304 > $O/dec_ab.log 2>&1
AB_TAG=old XZB200_LIB=xz_b200/libxzb200_decgen.so python tests/ab_probe.py T,6,33554432,4194304 E,6,33554432,4194304 R,3,33554432,4194304 >> $O/dec_ab.log 2>&1
grep -c "^OK" $O/probe_q.log; grep -v "^OK" $O/probe_q.log | tail -5
grep "DPPROF chain\|DPPROF nodes\|OK\|MISM" $O/prof_a.log $O/prof_b.log $O/prof_c.log | cut -c1-300
cat $O/dec_ab.log
