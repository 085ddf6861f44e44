#!/bin/bash
# BASELINE parity configurations C3 (mDBC + laminar+SPS) and C4 (piston + floating box) at scale on 2 GPUs (slabs along x)
mkdir -p gpurun_out
for cfg in c3 c4; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/run_config.py $cfg --npgpu ${1:-4000000} --steps 10 --warmup 3 > gpurun_out/r01b_config_$cfg.json 2> gpurun_out/r01b_config_$cfg.err
  echo "$cfg rc=$?"; cat gpurun_out/r01b_config_$cfg.json; tail -3 gpurun_out/r01b_config_$cfg.err | cut -c1-300
done
