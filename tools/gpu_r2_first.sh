#!/bin/bash
# First GPU call of the next round: (1) the whole GPU suite and the bench line with the shipped build, (2) A/B of the two prepared
# experiments of forces_pair2.cu, each rebuilt on the box (7 s), checked for parity of variant 5 on every fixture, then benched:
#   P2_VNOW: the exact-test branch of a round also waits for the velrho gathers -> the four gathers of a round are in flight together
#   P2_CHK2: leaner FP16 pre-filter (no index clamp, immediate mask bits, one x-block select per pair): 13 instead of 22 instructions per pair
# and one full ncu capture of the last build for the source page (tools/ncu_source_regions.py).
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r02_gpu_tests.log; cat gpurun_out/r02_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1_base.json 2> gpurun_out/r02_bench_n1_base.err; head -c 330 gpurun_out/r02_bench_n1_base.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_base.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02_launches_base.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_base.csv 2>/dev/null | head -12
python tools/gpu_r2_checks.py 2>&1 | tail -8 > gpurun_out/r02_host_checks.log; cat gpurun_out/r02_host_checks.log
k=0
for flags in "-DP2_VNOW=1" "-DP2_CHK2=1" "-DP2_VNOW=1 -DP2_CHK2=1"; do
  k=$((k+1))
  touch dualsphysics_b200/csrc/forces_pair2.cu
  make -C dualsphysics_b200/csrc B2S_NVCC_EXTRA="$flags" 2>&1 | tail -1 | cut -c1-80
  echo "== experiment $k: $flags"
  python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -3 > gpurun_out/r02_x${k}_parity.log; cat gpurun_out/r02_x${k}_parity.log
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_n1_x${k}.json 2> gpurun_out/r02_bench_n1_x${k}.err; head -c 330 gpurun_out/r02_bench_n1_x${k}.json; echo
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_forces_fluid_pair2 -s 2 -c 1 -f -o gpurun_out/r02_forces_pair2_x3_10M python tools/profile_forces.py 10000000 5 3 > gpurun_out/r02_prof_x3.log 2>&1
tail -1 gpurun_out/r02_prof_x3.log
