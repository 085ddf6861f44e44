#!/bin/bash
# Two-GPU job (gpurun --gpus 2): NCCL tests of the C++ slab step, then the 2-GPU bench.  Usage: bash tools/gpu_job_n2.sh TAG [np]
TAG=$1; NP=${2:-10000000}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dist.py tests/test_gpu_slab.py -q -m gpu -s 2>&1 | grep -E "^\[slab|SLAB_NCCL|passed|failed|Error|error|assert|^E " | head -80 > gpurun_out/${TAG}_dist_tests.log; tail -25 gpurun_out/${TAG}_dist_tests.log
B2S_BENCH_NP=$NP timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; tail -3 gpurun_out/${TAG}_bench_n2.err
python -c "import json; d=json.loads(open('gpurun_out/${TAG}_bench_n2.json').read().strip().splitlines()[-1]); print('N=2 ms/step', d['ms_per_step'], 'value', d['value'], 'np', d['config']['np_per_gpu'], 'comm', d['config'].get('comm'))"
