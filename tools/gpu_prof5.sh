#!/bin/bash
# usage: gpu_prof5.sh <lib.so relative path> <variant> <tag>
mkdir -p gpurun_out
export B2S_LIB_PATH=$PWD/$1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_forces_fluid -s 1 -c 1 -f -o gpurun_out/r01_forces_$3 python tools/profile_forces.py 2000000 $2 20 > gpurun_out/r01_prof_$3.log 2>&1
tail -1 gpurun_out/r01_prof_$3.log
