#!/bin/bash
# full GPU regression + a short N=1 bench (device-timed value only) after a kernel change
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r01b_gpu_tests.log; cat gpurun_out/r01b_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r01b_bench_quick.json 2> gpurun_out/r01b_bench_quick.err
python -c "import json;d=json.load(open('gpurun_out/r01b_bench_quick.json'));print(d['value'],d['ms_per_step'],d['roofline']['ms_per_launch'],d['gpu_launches'])"
