#!/bin/bash
# r02 call B: first run of interaction variant 6 (forces_tile.cu): memcheck on a small case, parity, bench, microbench.
mkdir -p gpurun_out
./tools/microbench > gpurun_out/r02b_microbench.txt 2>&1; cat gpurun_out/r02b_microbench.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "tile_variant" -s 2>&1 | grep -E "variant 6 vs 0|passed|failed|Error|error" | head -60 > gpurun_out/r02b_tile_vs_gather.log; cat gpurun_out/r02b_tile_vs_gather.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py 2 > gpurun_out/r02b_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02b_memcheck.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "golden and 6" 2>&1 | tail -5 > gpurun_out/r02b_golden_v6.log; cat gpurun_out/r02b_golden_v6.log
for v in 6 5; do
timeout 600 python bench.py --steps 10 --warmup 3 --variant $v --no-cpu-baseline --no-e2e > gpurun_out/r02b_bench_n1_v$v.json 2> gpurun_out/r02b_bench_n1_v$v.err; tail -3 gpurun_out/r02b_bench_n1_v$v.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02b_bench_n1_v$v.json"))
print("variant $v", d["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"])
PY
done
