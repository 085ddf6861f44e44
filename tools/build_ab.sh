#!/bin/bash
# Builds an A/B copy of the library here (no GPU needed): forces_pair2.cu recompiled with extra flags, everything else as built.
#   bash tools/build_ab.sh NAME "-DP2_MINB=8 ..."   ->  dualsphysics_b200/lib/ab/NAME/libb200sph.so   (git-ignored, travels with gpurun)
# On the box: B2S_LIB_PATH=$PWD/dualsphysics_b200/lib/ab/NAME/libb200sph.so python bench.py ...   (gpu_job.sh step ab=NAME1,NAME2)
set -e
NAME=$1; FLAGS=$2; SRC=${3:-forces_pair2}
cd "$(dirname "$0")/../dualsphysics_b200/csrc"
OUT=../lib/ab/$NAME; mkdir -p $OUT
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr $FLAGS -c $SRC.cu -o $OUT/$SRC.o 2> $OUT/$SRC.ptxas.log
OBJS=$(ls ../lib/obj/*.o | grep -v "/$SRC.o" | grep -v forces_defer)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libb200sph.so $OBJS $OUT/$SRC.o -lcudart -ldl
grep -A2 "k_forces_fluid_pair2ILi2ELi1ELi2ELb0ELb0" $OUT/$SRC.ptxas.log | tail -2
