#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error|error|assert" gpurun_out/pytest_gpu.log | cut -c1-600 | tail -20
run(){ # name variant extra
  B2S_BENCH_EXTRA_OPTS="$3" timeout 600 python bench.py --steps 5 --warmup 3 --variant $2 --no-cpu-baseline --no-e2e > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$1.json"))
    r=d["roofline"]
    print("$1: value %.4g ms/step %.2f  kernel ms %.3f frac %.3f share %.2f launches %d"%(d["value"],d["ms_per_step"],r["ms_per_launch"],r["frac"],r["share_of_step"],d["gpu_launches"]))
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/bench_$1.err").read()[-1500:])
PY
}
run v2 2 ""
run v0half 0 " -cellmode:half"
