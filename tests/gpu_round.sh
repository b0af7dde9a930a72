#!/bin/bash
O=gpurun_out
nvidia-smi -L | head -4
echo "== bench at 2 GPUs (torchrun)"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 3 --no-cpu-baseline > $O/bench_r02_2gpu.json 2> $O/bench_r02_2gpu.err; tail -c 600 $O/bench_r02_2gpu.json; tail -2 $O/bench_r02_2gpu.err | cut -c1-200
echo "== in-library fan-out over the GPUs of the box + LZMA_RUN-only decode"; timeout 200 python -m pytest tests/test_gpu_hybrid.py tests/test_gpu_lzma_api.py -q -x -k "fans_out or default_threads or lzma_run_only" > $O/pytest_2gpu.log 2>&1; tail -3 $O/pytest_2gpu.log | cut -c1-200
