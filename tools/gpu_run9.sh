#!/bin/bash
# GPU regression: all -m gpu tests (incl. shifting / damping / periodic fixtures), failures listed
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -k "${1:-}" 2>&1 | tail -60 > gpurun_out/gpu_tests9.log; tail -40 gpurun_out/gpu_tests9.log
