#!/bin/bash
# Scale job (gpurun --gpus N): the bench on N GPUs (10M particles per GPU, e2e leg included, phase profile), then one BASELINE
# configuration at its stated size.  Usage: bash tools/gpu_job_scale.sh TAG NGPUS [c3|c4|none] [np per GPU of the configuration]
TAG=$1; NG=$2; CFG=${3:-none}; NPC=${4:-12500000}
mkdir -p gpurun_out
B2S_SLAB_PROFILE=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n${NG}.json 2> gpurun_out/${TAG}_bench_n${NG}.err
grep -E -A9 "slab phase" gpurun_out/${TAG}_bench_n${NG}.err | head -12; grep -E "Error|error" gpurun_out/${TAG}_bench_n${NG}.err | tail -3
python -c "import json; d=json.loads(open('gpurun_out/${TAG}_bench_n${NG}.json').read().strip().splitlines()[-1]); print('N=${NG} ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', (d.get('e2e') or {}).get('value'), 'np', d['config']['np_per_gpu'], 'comm', {k:v for k,v in (d['config'].get('comm') or {}).items() if k!='note'})"
if [ "$CFG" != none ]; then
  B2S_SLAB_PROFILE=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29519 tools/run_config.py $CFG --npgpu $NPC --steps 10 --warmup 3 > gpurun_out/${TAG}_config_${CFG}_n${NG}.json 2> gpurun_out/${TAG}_config_${CFG}_n${NG}.err
  grep -E "Error|error" gpurun_out/${TAG}_config_${CFG}_n${NG}.err | tail -3; cat gpurun_out/${TAG}_config_${CFG}_n${NG}.json | cut -c1-1500
fi
