#!/bin/bash
# GPU session AC (round 2): suite, step times and bench lines with the one-thread-per-variant kernel and programmatic dependent
# launch in all three fused steps.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r2ac_tests.log; grep -E "passed|failed|FAILED" gpurun_out/r2ac_tests.log | tail -5
for w in c5_genome_scale c5_quarter; do timeout 200 python tools/quick_phase.py --tag main --workload $w >> gpurun_out/r2ac_phase.jsonl 2>> gpurun_out/r2ac_phase.err; done; cat gpurun_out/r2ac_phase.jsonl
timeout 300 python profiles/config_timings.py > gpurun_out/r2ac_config_timings.jsonl 2> gpurun_out/r2ac_config_timings.err; cat gpurun_out/r2ac_config_timings.jsonl
timeout 300 python profiles/survival_steady.py 300 > gpurun_out/r2ac_survival_steady.log 2>&1; cat gpurun_out/r2ac_survival_steady.log
timeout 600 python bench.py > gpurun_out/r2ac_bench_c5.json 2> gpurun_out/r2ac_bench_c5.err; tail -c 300 gpurun_out/r2ac_bench_c5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ac_bench_c5.json').read().strip().splitlines()[-1]); r=d['roofline']
print({k:d[k] for k in ('ms_per_step','value')}, 'e2e', d['e2e']['value'], d['e2e'].get('ms_total_of_each_run'), 'guide/alpha/var', r['ms_per_launch'], r['ms_per_launch_alpha_kernel'], r['ms_per_launch_variant_kernel'], 'full', d['full_run']['seconds'])
PY
