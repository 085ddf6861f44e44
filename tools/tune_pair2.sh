#!/bin/bash
# Builds libb200sph variants with different forces_pair2.cu tuning macros into dualsphysics_b200/lib/tune/ (run here, no GPU).
# usage: tools/tune_pair2.sh "tag -DP2_MINB=5 -DP2_MASK=2 ..." ...
set -e
cd "$(dirname "$0")/../dualsphysics_b200/csrc"
make -j8 > /dev/null
mkdir -p ../lib/tune
for cfg in "$@"; do
  set -- $cfg
  tag=$1; shift
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v \
    "$@" -c forces_pair2.cu -o ../lib/tune/$tag.o 2>&1 | grep -A2 "pair2ILi2" | grep -E "registers|spill" | tr '\n' ' '
  objs=$(ls ../lib/obj/*.o | grep -v forces_pair2.o)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/tune/libb200sph_$tag.so $objs ../lib/tune/$tag.o -lcudart
  rm ../lib/tune/$tag.o
  echo " built $tag"
done
