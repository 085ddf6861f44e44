#!/bin/bash
# r02 call A: calibrated microbench, GPU suite + bench of the shipped build, A/B of the two prepared pair2 experiments.
mkdir -p gpurun_out
./tools/microbench > gpurun_out/r02_microbench.txt 2>&1; cat gpurun_out/r02_microbench.txt
python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > gpurun_out/r02_gpu_tests.log; cat gpurun_out/r02_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1_base.json 2> gpurun_out/r02_bench_n1_base.err; head -c 330 gpurun_out/r02_bench_n1_base.json; echo
k=0
for flags in "-DP2_VNOW=1" "-DP2_CHK2=1" "-DP2_VNOW=1 -DP2_CHK2=1"; do
  k=$((k+1))
  touch dualsphysics_b200/csrc/forces_pair2.cu
  make -C dualsphysics_b200/csrc B2S_NVCC_EXTRA="$flags" 2>&1 | tail -1 | cut -c1-80
  echo "== experiment $k: $flags"
  python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3 > gpurun_out/r02_x${k}_parity.log; cat gpurun_out/r02_x${k}_parity.log
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_n1_x${k}.json 2> gpurun_out/r02_bench_n1_x${k}.err; head -c 330 gpurun_out/r02_bench_n1_x${k}.json; echo
  python - <<PY
import json
d=json.load(open("gpurun_out/r02_bench_n1_x${k}.json"))
print("x${k}", d["ms_per_step"], d["roofline"]["ms_per_launch"])
PY
done
