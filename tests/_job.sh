timeout 900 python -m pytest tests/test_gpu_lzma_api.py -x -q -k "sequences" 2>&1 | tail -4
timeout 600 ncu --clock-control none --set full --import-source on -k regex:xzb_k_parse_warp -c 1 -o gpurun_out/r01_parse_fast2_R3 python tests/ab_probe.py R,3,2097152,262144 > /dev/null 2>&1
timeout 600 ncu --clock-control none --set full --import-source on -k regex:xzb_k_parse_warp -c 1 -o gpurun_out/r01_parse_fast2_T1 python tests/ab_probe.py T,1,2097152,262144 > /dev/null 2>&1
ls -la gpurun_out | tail -3
