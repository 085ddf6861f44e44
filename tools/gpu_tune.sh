#!/bin/bash
# times the fluid interaction kernel (variant 3) for every tuning build in dualsphysics_b200/lib/tune, plus variant 0/3 of the default build
mkdir -p gpurun_out; : > gpurun_out/tune.txt
NP=${1:-10000000}
echo "default v0: $(python tools/profile_forces.py $NP 0 10 | tail -1)" >> gpurun_out/tune.txt
echo "default v3: $(python tools/profile_forces.py $NP 3 10 | tail -1)" >> gpurun_out/tune.txt
for f in dualsphysics_b200/lib/tune/*.so; do
  echo "$(basename $f): $(B2S_LIB_PATH=$PWD/$f python tools/profile_forces.py $NP 3 10 | tail -1)" >> gpurun_out/tune.txt
done
cat gpurun_out/tune.txt
