#!/bin/bash
# GPU session Z (round 2): final single-GPU validation -- suite, smoke(), the default bench line, the reference arm, c4.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r2z_tests.log; grep -E "passed|failed|FAILED" gpurun_out/r2z_tests.log | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -3 gpurun_out/r2z_smoke.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/r2z_bench_c5.json 2> gpurun_out/r2z_bench_c5.err; tail -c 300 gpurun_out/r2z_bench_c5.err
timeout 400 python bench.py --workload c4_survival > gpurun_out/r2z_bench_c4.json 2> gpurun_out/r2z_bench_c4.err; tail -c 300 gpurun_out/r2z_bench_c4.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_c5_reference_arm.json 2> gpurun_out/r2z_bench_c5_reference_arm.err; tail -c 300 gpurun_out/r2z_bench_c5_reference_arm.err
python - <<'PY'
import json
for f in ('gpurun_out/r2z_bench_c5.json','gpurun_out/r2z_bench_c4.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
    print(f, {k:d[k] for k in ('ms_per_step','value','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e'].get('ms_total_of_each_run'), 'guide/alpha/var', r['ms_per_launch'], r['ms_per_launch_alpha_kernel'], r['ms_per_launch_variant_kernel'], 'full', d['full_run']['seconds'], 'frac', r['frac'], 'traffic', r['traffic'], 'ceil', r['compute_roofline']['row_math_ceiling_ms'], r['compute_roofline']['frac'], r['compute_roofline']['issue_floor_ms'], 'cpu', d.get('cpu_baseline',{}).get('value'), d['clocks'])
d=json.loads(open('gpurun_out/r2z_bench_c5_reference_arm.json').read().strip().splitlines()[-1])
print('reference arm', d['value'], d['ms_per_step'], d['anomaly_mode_off'])
PY
