#!/bin/bash
O=gpurun_out
echo "== full gpu suite"; timeout 1000 python -m pytest tests -q -m gpu -x > $O/pytest_gpu_full.log 2>&1; tail -6 $O/pytest_gpu_full.log | cut -c1-250
echo "== smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2 | cut -c1-300
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 > $O/bench_r02_a.json 2> $O/bench_r02_a.err; tail -c 3000 $O/bench_r02_a.json; tail -3 $O/bench_r02_a.err
