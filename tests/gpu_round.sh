#!/bin/bash
O=gpurun_out
echo "== full gpu suite"; timeout 900 python -m pytest tests -q -m gpu -x > $O/pytest_gpu_full.log 2>&1; tail -4 $O/pytest_gpu_full.log | cut -c1-250
echo "== ncu launch list"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-lzma-code > $O/r02_launches_bench.json 2> $O/r02_launches.err; wc -l $O/r02_launches.csv; tail -c 300 $O/r02_launches_bench.json
