#!/bin/bash
# First GPU call of the next round: (1) the whole GPU suite and the bench line with the shipped build, (2) A/B of the P2_VNOW
# experiment (forces_pair2.cu: the exact-test branch of a round also waits for the velrho gathers, so the four gathers of a round
# are in flight together): rebuilt on the box, parity of variant 5 on every fixture, bench line, kernel time from the ncu launch list.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r02_gpu_tests.log; cat gpurun_out/r02_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1_base.json 2> gpurun_out/r02_bench_n1_base.err; head -c 330 gpurun_out/r02_bench_n1_base.json; echo
touch dualsphysics_b200/csrc/forces_pair2.cu
make -C dualsphysics_b200/csrc EXTRA=-DP2_VNOW=1 2>&1 | tail -1
python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -3 > gpurun_out/r02_vnow_parity.log; cat gpurun_out/r02_vnow_parity.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1_vnow.json 2> gpurun_out/r02_bench_n1_vnow.err; head -c 330 gpurun_out/r02_bench_n1_vnow.json; echo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_forces_fluid_pair2 -s 2 -c 1 -f -o gpurun_out/r02_forces_pair2_vnow_10M python tools/profile_forces.py 10000000 5 3 > gpurun_out/r02_prof_vnow.log 2>&1
tail -1 gpurun_out/r02_prof_vnow.log
