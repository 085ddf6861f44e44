#!/bin/bash
# round-1 final evidence at N=1: GPU tests, bench line, launch list, full ncu capture of the dominant kernel at the bench size
mkdir -p gpurun_out
echo "(GPU tests: see r01b_gpu_tests.log of the previous call)"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r01b_bench_n1.json 2> gpurun_out/r01b_bench_n1.err; head -c 400 gpurun_out/r01b_bench_n1.json; echo
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01b_bench_ref.json 2> gpurun_out/r01b_bench_ref.err; head -c 300 gpurun_out/r01b_bench_ref.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r01b_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_forces_fluid_pair2 -s 2 -c 1 -f -o gpurun_out/r01b_forces_pair2_10M python tools/profile_forces.py 10000000 4 3 > gpurun_out/r01b_prof_pair2_10M.log 2>&1
tail -1 gpurun_out/r01b_prof_pair2_10M.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
