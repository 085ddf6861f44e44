#!/bin/bash
mkdir -p gpurun_out
python tools/debug_fx.py ver_cubic_art_ddt1 > gpurun_out/debug_fx.txt 2>&1
python tools/debug_fx.py sym_wend_art_ddt2 >> gpurun_out/debug_fx.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
cat gpurun_out/debug_fx.txt
grep -E "^\[parity|passed|failed|FAILED|Error|error" gpurun_out/pytest_gpu.log | cut -c1-1500 | tail -40
