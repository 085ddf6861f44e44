#!/bin/bash
# final evidence with interaction variant 5 (mirror-row pairing) as the bench default: GPU tests, bench line, full ncu capture
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r01c_gpu_tests.log; cat gpurun_out/r01c_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r01c_bench_n1.json 2> gpurun_out/r01c_bench_n1.err; head -c 330 gpurun_out/r01c_bench_n1.json; echo
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_forces_fluid_pair2 -s 2 -c 1 -f -o gpurun_out/r01c_forces_pair2_v5_10M python tools/profile_forces.py 10000000 5 3 > gpurun_out/r01c_prof_v5.log 2>&1
tail -1 gpurun_out/r01c_prof_v5.log
