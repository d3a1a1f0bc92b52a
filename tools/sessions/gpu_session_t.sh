#!/bin/bash
# GPU session T (round 2): the suite, fp32 error report and step times after the MUFU float tail regimes; bench lines.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r2t_tests.log
grep -E "passed|failed|FAILED" gpurun_out/r2t_tests.log | tail
timeout 200 python tools/fp32_error_report.py > gpurun_out/r2t_fp32_errors.json 2> gpurun_out/r2t_fp32_errors.err
timeout 200 python tools/quick_phase.py --tag main > gpurun_out/r2t_phase.json 2> gpurun_out/r2t_phase.err; cat gpurun_out/r2t_phase.json
timeout 300 python profiles/survival_steady.py 300 > gpurun_out/r2t_survival_steady.log 2>&1; cat gpurun_out/r2t_survival_steady.log
timeout 300 python profiles/config_timings.py > gpurun_out/r2t_config_timings.jsonl 2> gpurun_out/r2t_config_timings.err; cat gpurun_out/r2t_config_timings.jsonl
timeout 600 python bench.py > gpurun_out/r2t_bench_c5.json 2> gpurun_out/r2t_bench_c5.err; tail -c 300 gpurun_out/r2t_bench_c5.err
timeout 400 python bench.py --workload c4_survival > gpurun_out/r2t_bench_c4.json 2> gpurun_out/r2t_bench_c4.err; tail -c 300 gpurun_out/r2t_bench_c4.err
python - <<'PY'
import json
for f in ('gpurun_out/r2t_bench_c5.json','gpurun_out/r2t_bench_c4.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
    print(f, {k:d[k] for k in ('ms_per_step','value')}, 'e2e', d['e2e']['value'], d['e2e'].get('ms_total'), 'guide/alpha/var', r['ms_per_launch'], r['ms_per_launch_alpha_kernel'], r['ms_per_launch_variant_kernel'], 'full', d['full_run']['seconds'], 'frac', r['frac'], 'ceil', r['compute_roofline']['row_math_ceiling_ms'], r['compute_roofline']['frac'], 'cpu', d.get('cpu_baseline',{}).get('value'))
PY
