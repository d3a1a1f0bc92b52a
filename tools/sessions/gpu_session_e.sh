#!/bin/bash
# GPU session E (round 2): suite, fp32 errors (double tail regimes), split vs single-kernel guide step, survival kernel profile.
mkdir -p gpurun_out
python -m crispr_bean_b200.build --force > gpurun_out/r2e_build.log 2>&1 || { tail -30 gpurun_out/r2e_build.log; exit 1; }
timeout 400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -120 > gpurun_out/r2e_tests.log
timeout 200 python tools/fp32_error_report.py > gpurun_out/r2e_fp32_errors.json 2> gpurun_out/r2e_fp32_errors.err
timeout 200 python tools/quick_phase.py --tag split > gpurun_out/r2e_phase_split.json 2> gpurun_out/r2e_phase_split.err
timeout 200 python tools/quick_phase.py --tag nosplit --no-split > gpurun_out/r2e_phase_nosplit.json 2> gpurun_out/r2e_phase_nosplit.err
timeout 300 python profiles/survival_steady.py 300 > gpurun_out/r2e_survival_steady.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:surv_guide_kernel -s 200 -c 1 -f -o gpurun_out/prof_r2e_surv_guide python profiles/survival_steady.py 300 > gpurun_out/r2e_ncu_surv.log 2>&1
grep -E "passed|failed|FAILED" gpurun_out/r2e_tests.log | tail
cat gpurun_out/r2e_phase_split.json gpurun_out/r2e_phase_nosplit.json gpurun_out/r2e_survival_steady.log
