#!/bin/bash
# Run under gpurun (one GPU).  Writes raw ncu output to gpurun_out/; summaries are made offline
# with profiles/ncu_lines.py / profiles/launch_summary.py and committed under profiles/.
# (ncu injects into the process, which makes the library run the match finder and the parser
# back to back instead of side by side -- see DESIGN.md section 4.)
set -x
O=gpurun_out
N="ncu --clock-control none"
# 1. launch list of one encode (device + e2e leg) + decode pass: 8 x 16 MiB blocks at -6
$N --metrics gpu__time_duration.sum --csv --log-file $O/r01_launches.csv \
   python bench.py --steps 1 --warmup 0 --size 134217728 --no-cpu-baseline --verify-blocks 0 > $O/r01_launches_bench.json 2> $O/r01_launches.err
# 2. full captures of the dominant kernels
$N --set full --import-source on -k regex:xzb_k_bt -c 1 -o $O/r01_bt \
   python bench.py --steps 1 --warmup 0 --size 134217728 --no-cpu-baseline --verify-blocks 0 > /dev/null 2>&1
$N --set full --import-source on -k regex:xzb_k_parse_warp -c 1 -o $O/r01_parse \
   python tests/gpu_probe.py T,6,2097152,262144 > /dev/null 2>&1
$N --set full --import-source on -k regex:xzb_k_hc -c 1 -o $O/r01_hc \
   python bench.py --preset 3 --kind R --steps 1 --warmup 0 --size 268435456 --no-cpu-baseline --verify-blocks 0 > /dev/null 2>&1
$N --set full --import-source on -k regex:xzb_k_decode -c 1 -o $O/r01_decode \
   python tests/gpu_probe.py T,6,2097152,262144 > /dev/null 2>&1
# fast mode (presets 0-3): parser kernel with the decision warp + coding warp
$N --set full --import-source on -k regex:xzb_k_parse_warp -c 1 -o $O/r01_parse_fast2_R3 python tests/ab_probe.py R,3,2097152,262144 > /dev/null 2>&1
$N --set full --import-source on -k regex:xzb_k_parse_warp -c 1 -o $O/r01_parse_fast2_T1 python tests/ab_probe.py T,1,2097152,262144 > /dev/null 2>&1
# 3. the HC4 / incompressible configuration as a plain bench line (configs[4] shape on one GPU: 1 GiB of it)
python bench.py --preset 3 --kind R --steps 1 --warmup 1 --no-cpu-baseline > $O/r01_bench_R3.json 2> $O/r01_bench_R3.err
ls -la $O
