#!/bin/bash
# GPU session AF (round 2): the suite and the bench line with the data-only row constants computed by bean_row_const (one pass).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r2af_tests.log; grep -E "passed|failed|FAILED" gpurun_out/r2af_tests.log | tail -5
timeout 200 python tools/e2e_breakdown.py > gpurun_out/r2af_e2e_breakdown.json 2> gpurun_out/r2af_e2e_breakdown.err; cat gpurun_out/r2af_e2e_breakdown.json
timeout 600 python bench.py > gpurun_out/r2af_bench_c5.json 2> gpurun_out/r2af_bench_c5.err; tail -c 300 gpurun_out/r2af_bench_c5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2af_bench_c5.json').read().strip().splitlines()[-1]); r=d['roofline']
print({k:d[k] for k in ('ms_per_step','value')}, 'e2e', d['e2e']['value'], d['e2e'].get('ms_total_of_each_run'), d['e2e'].get('ms_upload_and_setup'), 'full', d['full_run']['seconds'], 'final_loss', d['final_loss'])
PY
