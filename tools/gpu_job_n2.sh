#!/bin/bash
# Multi-GPU job (gpurun --gpus N): NCCL tests of the C++ slab step, then the N-GPU bench with the phase profile of the slab step.
# Usage: bash tools/gpu_job_n2.sh TAG [np per GPU] [ngpus] [tests|notests]
TAG=$1; NP=${2:-10000000}; NG=${3:-2}; T=${4:-tests}
mkdir -p gpurun_out
if [ "$T" = tests ]; then
  timeout 1200 python -m pytest tests/test_gpu_dist.py tests/test_gpu_slab.py -q -m gpu -s 2>&1 | grep -E "^\[slab|SLAB_NCCL|passed|failed|Error|error|assert|^E " | head -80 > gpurun_out/${TAG}_dist_tests.log; tail -8 gpurun_out/${TAG}_dist_tests.log
fi
B2S_SLAB_PROFILE=1 B2S_BENCH_NP=$NP timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 10 --warmup 3 ${BENCH_EXTRA---no-e2e} --no-cpu-baseline > gpurun_out/${TAG}_bench_n${NG}.json 2> gpurun_out/${TAG}_bench_n${NG}.err; grep -E -A9 "slab phase" gpurun_out/${TAG}_bench_n${NG}.err | head -12; grep -E "Error|error" gpurun_out/${TAG}_bench_n${NG}.err | tail -3
python -c "import json; d=json.loads(open('gpurun_out/${TAG}_bench_n${NG}.json').read().strip().splitlines()[-1]); print('N=${NG} ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', (d.get('e2e') or {}).get('value'), 'np', d['config']['np_per_gpu'], 'comm', {k:v for k,v in (d['config'].get('comm') or {}).items() if k!='note'})"
