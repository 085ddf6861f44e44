#!/bin/bash
# r01: slab tests + full GPU suite + single-GPU bench (new defaults: variant 0, -cellmode:half, event timing)
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -s 2>&1 | grep -a "slab \|passed\|failed\|Error\|error" | tail -30 > gpurun_out/gpu_tests6.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r01_bench_n1.json 2> gpurun_out/bench6.err
tail -c 3000 gpurun_out/r01_bench_n1.json; tail -5 gpurun_out/bench6.err; cat gpurun_out/gpu_tests6.log
