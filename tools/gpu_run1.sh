#!/bin/bash
# first GPU pass: parity tests, microbench, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 120 ./tools/microbench > gpurun_out/microbench.txt 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --variant 0 > gpurun_out/bench_v0.json 2> gpurun_out/bench_v0.err
tail -c 3000 gpurun_out/pytest_gpu.log
cat gpurun_out/microbench.txt
cat gpurun_out/bench_v0.json
tail -5 gpurun_out/bench_v0.err
