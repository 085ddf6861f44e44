#!/bin/bash
mkdir -p gpurun_out
V=${1:-3}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_forces_fluid -s 1 -c 1 -f -o gpurun_out/r01_forces_v$V python tools/profile_forces.py 2000000 $V 20 > gpurun_out/r01_prof_v$V.log 2>&1
tail -2 gpurun_out/r01_prof_v$V.log
