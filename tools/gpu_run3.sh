#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error|error|assert" gpurun_out/pytest_gpu.log | cut -c1-600 | tail -30
for v in 1 0; do
  timeout 600 python bench.py --steps 5 --warmup 3 --variant $v --no-cpu-baseline > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_v$v.json"))
    r=d["roofline"]
    print("variant $v: value %.4g ms/step %.2f  kernel ms %.3f frac %.3f share %.2f e2e %.4g launches %d clocks %s"%(d["value"],d["ms_per_step"],r["ms_per_launch"],r["frac"],r["share_of_step"],d["e2e"]["value"],d["gpu_launches"],d["clocks"]))
except Exception as e:
    print("variant $v failed", e); print(open("gpurun_out/bench_v$v.err").read()[-2000:])
PY
done
