#!/bin/bash
# GPU session N (round 2): suite + phase times after the per-guide near-mean coefficients of the alpha kernel; survival steady.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r2n_tests.log
grep -E "passed|failed|FAILED" gpurun_out/r2n_tests.log | tail
timeout 200 python tools/quick_phase.py --tag main > gpurun_out/r2n_phase.json 2> gpurun_out/r2n_phase.err; cat gpurun_out/r2n_phase.json
timeout 200 python tools/fp32_error_report.py > gpurun_out/r2n_fp32_errors.json 2> gpurun_out/r2n_fp32_errors.err; tail -c 1500 gpurun_out/r2n_fp32_errors.json
timeout 300 python profiles/survival_steady.py 300 > gpurun_out/r2n_survival_steady.log 2>&1; cat gpurun_out/r2n_survival_steady.log
du -sh gpurun_out
