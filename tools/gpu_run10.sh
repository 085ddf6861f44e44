#!/bin/bash
# 2-GPU regression after the next-row work: NCCL slab tests + a short N=2 bench line
mkdir -p gpurun_out
python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -3 > gpurun_out/r01b_gpu_dist.log; cat gpurun_out/r01b_gpu_dist.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r01b_bench_n2.json 2> gpurun_out/r01b_bench_n2.err
head -c 600 gpurun_out/r01b_bench_n2.json; echo; tail -3 gpurun_out/r01b_bench_n2.err
