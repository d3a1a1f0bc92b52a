#!/bin/bash
# GPU session L (round 2): suite after the FP64-divide slow-path fix of the tiling step, A/B of guide-kernel (FSET, row-mask bits)
# and alpha-kernel (occupancy) builds, tiling step time, default bench line with the e2e breakdown.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r2l_tests.log
grep -E "passed|failed|FAILED" gpurun_out/r2l_tests.log | tail
timeout 120 python profiles/tiling_steady.py 300 > gpurun_out/r2l_tiling_steady.log 2>&1; cat gpurun_out/r2l_tiling_steady.log
timeout 200 python tools/quick_phase.py --tag main >> gpurun_out/r2l_variants.jsonl 2>> gpurun_out/r2l_variants.err
for v in fset rmask both2 alpha8 alpha5; do
  BEAN_B200_LIB=$PWD/crispr_bean_b200/variants/libbean_b200_$v.so timeout 200 python tools/quick_phase.py --tag $v >> gpurun_out/r2l_variants.jsonl 2>> gpurun_out/r2l_variants.err
done
cat gpurun_out/r2l_variants.jsonl
timeout 300 python profiles/config_timings.py > gpurun_out/r2l_config_timings.jsonl 2> gpurun_out/r2l_config_timings.err; cat gpurun_out/r2l_config_timings.jsonl
timeout 600 python bench.py > gpurun_out/r2l_bench_c5.json 2> gpurun_out/r2l_bench_c5.err; tail -c 300 gpurun_out/r2l_bench_c5.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2l_bench_c5.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('ms_per_step','value')}, d['e2e'])
PY
du -sh gpurun_out
