#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s 2>&1 | grep -a "variant 4\|passed\|failed\|Error\|assert" | tail -14 > gpurun_out/gpu_tests8.log
cat gpurun_out/gpu_tests8.log
for v in 0 4; do echo "v$v: $(python tools/profile_forces.py 10000000 $v 10 | tail -1)"; done | tee gpurun_out/time8.txt
