#!/bin/bash
# One parametrised GPU job (replaces the numbered one-shot scripts of round 1).  Usage (inside gpurun):
#   bash tools/gpu_job.sh TAG [steps...]      steps: see the case labels below (suite, bench5, ncu5, launches, refgpu, parity_report, memcheck ...)
# Outputs go to gpurun_out/TAG_*.
TAG=$1; shift
mkdir -p gpurun_out
for step in "$@"; do
case $step in
  build=*) flags=$(echo "${step#build=}" | tr ',' ' '); touch dualsphysics_b200/csrc/forces_tile.cu dualsphysics_b200/csrc/forces_pair2.cu; make -C dualsphysics_b200/csrc B2S_NVCC_EXTRA="$flags" 2>&1 | grep -E "error|rror:" ; echo "== rebuilt with: $flags"; TAG=${TAG}x;;
  ab=*) for n in $(echo "${step#ab=}" | tr ',' ' '); do B2S_LIB_PATH=$PWD/dualsphysics_b200/lib/ab/$n/libb200sph.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ab_$n.json 2> gpurun_out/${TAG}_ab_$n.err; python -c "import json; d=json.load(open('gpurun_out/${TAG}_ab_$n.json')); print('$n: ms/step', round(d['ms_per_step'],3), 'ms/interaction', round(d['roofline']['ms_per_launch'],3), 'parity', d.get('parity'))"; done | tee gpurun_out/${TAG}_ab.txt;;   # libraries prebuilt by tools/build_ab.sh
  microbench) ./tools/microbench > gpurun_out/${TAG}_microbench.txt 2>&1; cat gpurun_out/${TAG}_microbench.txt;;
  pack_parity) timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "packed_variant or (golden and 5)" -s 2>&1 | grep -E "variant 5 vs 0|passed|failed|Error|error|assert" | head -40 > gpurun_out/${TAG}_pack_vs_gather.log; tail -12 gpurun_out/${TAG}_pack_vs_gather.log;;
  tile_parity) timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "tile_variant" -s 2>&1 | grep -E "variant 6 vs 0|passed|failed|Error|error|assert" | head -60 > gpurun_out/${TAG}_tile_vs_gather.log; tail -3 gpurun_out/${TAG}_tile_vs_gather.log;;
  golden6) timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "golden and 6" 2>&1 | tail -5 > gpurun_out/${TAG}_golden_v6.log; cat gpurun_out/${TAG}_golden_v6.log;;
  memcheck) timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py 2 > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/${TAG}_memcheck.log;;
  racecheck) timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py 1 > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/${TAG}_racecheck.log;;
  bench[0-9]) v=${step#bench}
    timeout 600 python bench.py --steps 10 --warmup 3 --variant $v --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_n1_v$v.json 2> gpurun_out/${TAG}_bench_n1_v$v.err; tail -2 gpurun_out/${TAG}_bench_n1_v$v.err
    python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench_n1_v$v.json')); print('variant $v ms/step', d['ms_per_step'], 'ms/interaction', d['roofline']['ms_per_launch'], 'frac', d['roofline']['frac'])";;
  benchfull6) B2S_BENCH_EXTRA_OPTS=' -cellmode:full' timeout 600 python bench.py --steps 10 --warmup 3 --variant 6 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_n1_v6_fullcells.json 2> gpurun_out/${TAG}_bench_n1_v6_fullcells.err; tail -2 gpurun_out/${TAG}_bench_n1_v6_fullcells.err
    python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench_n1_v6_fullcells.json')); print('variant 6 full cells ms/step', d['ms_per_step'], 'ms/interaction', d['roofline']['ms_per_launch'])";;
  bench25) B2S_BENCH_NP=25000000 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_n1_25M.json 2> gpurun_out/${TAG}_bench_n1_25M.err; tail -2 gpurun_out/${TAG}_bench_n1_25M.err; python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench_n1_25M.json')); print('25M: ms/step', d['ms_per_step'], 'value', d['value'], 'ms/interaction', d['roofline']['ms_per_launch'], 'frac', d['roofline']['frac'])";;
  fullbench) timeout 1200 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -2 gpurun_out/${TAG}_bench_n1.err; head -c 400 gpurun_out/${TAG}_bench_n1.json; echo;;
  ncu[0-9]) v=${step#ncu}
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_forces_fluid -s 2 -c 1 -f -o gpurun_out/${TAG}_forces_v${v}_10M python tools/profile_forces.py 10000000 $v 3 > gpurun_out/${TAG}_prof_v$v.log 2>&1; tail -1 gpurun_out/${TAG}_prof_v$v.log
    python tools/ncu_keys.py gpurun_out/${TAG}_forces_v${v}_10M.ncu-rep > gpurun_out/${TAG}_ncu_forces_v${v}_10M.txt 2>&1; python tools/ncu_source_regions.py gpurun_out/${TAG}_forces_v${v}_10M.ncu-rep > gpurun_out/${TAG}_source_regions_v${v}.txt 2>&1
    [ $(stat -c %s gpurun_out/${TAG}_forces_v${v}_10M.ncu-rep) -gt 30000000 ] && rm -f gpurun_out/${TAG}_forces_v${v}_10M.ncu-rep;;
  launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_launches.log 2>&1; python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv 2>/dev/null | head -14;;
  parity_report) rm -f gpurun_out/${TAG}_parity.jsonl; B2S_PARITY_REPORT=$PWD/gpurun_out/${TAG}_parity.jsonl timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py -q -m gpu -k "golden or large" 2>&1 | tail -4 > gpurun_out/${TAG}_parity_tests.log; cat gpurun_out/${TAG}_parity_tests.log; wc -l gpurun_out/${TAG}_parity.jsonl;;
  ncustep=*) kind=${step#ncustep=}; npk=10000000; [ "$kind" != bench ] && npk=4000000
    timeout 1500 ncu --set full --clock-control none --profile-from-start off --metrics sm__sass_thread_inst_executed_op_fadd_pred_on.sum,sm__sass_thread_inst_executed_op_fmul_pred_on.sum,sm__sass_thread_inst_executed_op_ffma_pred_on.sum -f -o gpurun_out/${TAG}_step_${kind} python tools/profile_step.py $kind $npk 30 > gpurun_out/${TAG}_step_${kind}.log 2>&1; tail -1 gpurun_out/${TAG}_step_${kind}.log
    python tools/ncu_step_summary.py gpurun_out/${TAG}_step_${kind}.ncu-rep > gpurun_out/${TAG}_ncu_step_${kind}.md 2>&1; rm -f gpurun_out/${TAG}_step_${kind}.ncu-rep   # the report of a whole step exceeds what a GPU job may return: the summary travels
    head -14 gpurun_out/${TAG}_ncu_step_${kind}.md | cut -c1-330;;
  radix_ab) for m in 1 4 8 16; do B2S_RADIX_PRE=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_radix_$m.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/${TAG}_radix_$m.json')); print('B2S_RADIX_PRE=$m ms/step', round(d['ms_per_step'],3), 'ms/interaction', round(d['roofline']['ms_per_launch'],3))"; done | tee gpurun_out/${TAG}_radix_ab.txt;;
  reorder_ab) for m in ${REORDER_MODES:-10 20 40}; do B2S_REORDER=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > gpurun_out/${TAG}_reorder_$m.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/${TAG}_reorder_$m.json')); print('B2S_REORDER=$m ms/step', round(d['ms_per_step'],3), 'ms/interaction', round(d['roofline']['ms_per_launch'],3))"; done | tee gpurun_out/${TAG}_reorder_ab.txt;;
  launches_c3) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_launches_c3.csv python tools/run_config.py c3 --npgpu 4000000 --steps 2 --warmup 3 > gpurun_out/${TAG}_launches_c3.log 2>&1; python tools/summarize_launches.py gpurun_out/${TAG}_launches_c3.csv 2>/dev/null | head -12;;
  c3small) timeout 900 python tools/run_config.py c3 --npgpu 4000000 --steps 10 > gpurun_out/${TAG}_c3_4M.json 2> gpurun_out/${TAG}_c3_4M.err; tail -c 600 gpurun_out/${TAG}_c3_4M.json; echo;;
  refgpu) timeout 1500 python tools/ref_gpu_ab.py 10000000 20 > gpurun_out/${TAG}_ref_gpu_ab.json 2> gpurun_out/${TAG}_ref_gpu_ab.err; tail -3 gpurun_out/${TAG}_ref_gpu_ab.err; head -c 3000 gpurun_out/${TAG}_ref_gpu_ab.json; echo;;
  suite) python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/${TAG}_gpu_tests.log; cat gpurun_out/${TAG}_gpu_tests.log;;
  *) echo "unknown step $step";;
esac
done
