#!/bin/bash
# Builds libb200sph variants with different forces_defer.cu tuning macros into dualsphysics_b200/lib/tune/ (run here, no GPU).
# usage: tools/tune_defer.sh "MINB CHUNK CAP" ...
set -e
cd "$(dirname "$0")/../dualsphysics_b200/csrc"
make -j8 > /dev/null
mkdir -p ../lib/tune
for cfg in "$@"; do
  set -- $cfg
  tag="m$1_c$2_q$3"
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
    -DD_MINB=$1 -DQ_CHUNK=$2 -DQ_CAP=$3 -c forces_defer.cu -o ../lib/tune/fd_$tag.o
  objs=$(ls ../lib/obj/*.o | grep -v forces_defer.o)
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/tune/libb200sph_$tag.so $objs ../lib/tune/fd_$tag.o -lcudart
  rm ../lib/tune/fd_$tag.o
  echo built $tag
done
