O=gpurun_out
timeout 900 python bench.py > $O/bench_r01_final.json 2> $O/bench_r01_final.err; tail -c 400 $O/bench_r01_final.err
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file $O/r01_launches.csv python bench.py --steps 1 --warmup 0 --size 134217728 --no-cpu-baseline --verify-blocks 0 > $O/r01_launches_bench.json 2> $O/r01_launches.err
timeout 400 python bench.py --preset 3 --kind R --steps 1 --warmup 1 --no-cpu-baseline > $O/r01_bench_R3.json 2> $O/r01_bench_R3.err
AB_TAG=memcheck timeout 400 compute-sanitizer --tool memcheck python tests/ab_probe.py T,1,262144,65536 T,6,262144,65536 R,3,131072,65536 2>&1 | tail -6
python -c "
import json
for f in ('bench_r01_final','r01_bench_R3'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['value'], d['e2e']['value'], d['decode'], d['kernels_ms'], d.get('cpu_baseline'), d['parity'], d['roofline']['frac'])"
