O=gpurun_out
N="ncu --clock-control none"
timeout 300 $N --set full --import-source on -k regex:xzb_k_parse_warp -c 1 -o $O/r01f_parse python tests/ab_probe.py T,6,2097152,262144 > /dev/null 2>&1
timeout 300 $N --set full --import-source on -k regex:xzb_k_decode -c 1 -o $O/r01f_decode python tests/ab_probe.py T,6,2097152,262144 > /dev/null 2>&1
timeout 300 $N --set full --import-source on -k regex:xzb_k_bt -c 1 -o $O/r01f_bt python tests/ab_probe.py T,6,33554432,4194304 > /dev/null 2>&1
ls -la $O/r01f_*
