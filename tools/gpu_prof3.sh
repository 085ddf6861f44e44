#!/bin/bash
# round-1 evidence: launch list of the bench command + full ncu capture of the dominant kernel (both cell modes)
mkdir -p gpurun_out
export B2S_BENCH_EXTRA_OPTS=" -cellmode:half"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_half.csv python bench.py --steps 2 --warmup 3 --variant 0 --no-cpu-baseline --no-e2e > gpurun_out/r01_launches_half.log 2>&1
export B2S_BENCH_EXTRA_OPTS=""
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_full.csv python bench.py --steps 2 --warmup 3 --variant 0 --no-cpu-baseline --no-e2e > gpurun_out/r01_launches_full.log 2>&1
B2S_BENCH_EXTRA_OPTS=" -cellmode:half" timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_forces_fluid_gather -s 1 -c 1 -o gpurun_out/r01_forces_gather_half python tools/profile_forces.py 2000000 0 20 > gpurun_out/r01_prof_half.log 2>&1
B2S_BENCH_EXTRA_OPTS="" timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_forces_fluid_gather -s 1 -c 1 -o gpurun_out/r01_forces_gather_full python tools/profile_forces.py 2000000 0 20 > gpurun_out/r01_prof_full.log 2>&1
tail -2 gpurun_out/r01_prof_half.log gpurun_out/r01_prof_full.log
ls -la gpurun_out/*.ncu-rep
# plain (unprofiled) bench numbers with the default settings
B2S_BENCH_EXTRA_OPTS=" -cellmode:half" timeout 900 python bench.py --steps 10 --warmup 3 --variant 0 > gpurun_out/r01_bench_half.json 2> gpurun_out/r01_bench_half.err
cat gpurun_out/r01_bench_half.json | cut -c1-400
