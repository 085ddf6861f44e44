#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/gpu_tests7.log
cat gpurun_out/gpu_tests7.log
for v in 0 3; do
  timeout 600 python bench.py --steps 5 --warmup 3 --variant $v --no-cpu-baseline --no-e2e > gpurun_out/bench7_v$v.json 2> gpurun_out/bench7_v$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench7_v$v.json"))
    print("variant $v", d["value"], d["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"])
except Exception as e:
    print("variant $v failed", e); print(open("gpurun_out/bench7_v$v.err").read()[-2000:])
PY
done
