for v in w0p0 w1p0 w0p1; do XZB200_LIB=xz_b200/ab/libxzb200_$v.so timeout 300 python tests/ab_probe.py T,6,16777216,2097152 E,6,16777216,2097152 R,3,16777216,2097152 T,1,16777216,2097152; done
timeout 300 python tests/ab_probe.py T,6,16777216,2097152 E,6,16777216,2097152 R,3,16777216,2097152 T,1,16777216,2097152
timeout 600 ncu --clock-control none --set full --import-source on -k regex:xzb_k_parse_warp -c 1 -o gpurun_out/r01_parse_fast_R3 python tests/ab_probe.py R,3,2097152,262144 > /dev/null 2>&1
timeout 600 ncu --clock-control none --set full --import-source on -k regex:xzb_k_parse_warp -c 1 -o gpurun_out/r01_parse_fast_T1 python tests/ab_probe.py T,1,2097152,262144 > /dev/null 2>&1
timeout 500 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ov3.json 2> gpurun_out/bench_ov3.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_ov3.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['decode'], d['kernels_ms'], d['parity'])"; tail -3 gpurun_out/bench_ov3.err
ls -la gpurun_out | tail -5
