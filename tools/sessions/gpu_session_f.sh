#!/bin/bash
# GPU session F (round 2): whole suite (incl. the 2000-step end-to-end check), phase times, ncu captures (guide, alpha, survival
# guide), default bench.
mkdir -p gpurun_out
python -m crispr_bean_b200.build --force > gpurun_out/r2f_build.log 2>&1 || { tail -30 gpurun_out/r2f_build.log; exit 1; }
timeout 600 python -m pytest tests -m gpu -q --tb=short --durations=8 2>&1 | tail -80 > gpurun_out/r2f_tests.log
timeout 200 python tools/fp32_error_report.py > gpurun_out/r2f_fp32_errors.json 2> gpurun_out/r2f_fp32_errors.err
timeout 200 python tools/quick_phase.py --tag final > gpurun_out/r2f_phase.json 2> gpurun_out/r2f_phase.err
timeout 600 bash profiles/capture.sh r2f > gpurun_out/r2f_capture.log 2>&1
timeout 300 python profiles/survival_steady.py 300 > gpurun_out/r2f_survival_steady.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:surv_guide_kernel -s 200 -c 1 -f -o gpurun_out/prof_r2f_surv_guide python profiles/survival_steady.py 300 > gpurun_out/r2f_ncu_surv.log 2>&1
grep -E "passed|failed|FAILED|^[0-9.]+s " gpurun_out/r2f_tests.log | tail -20
cat gpurun_out/r2f_phase.json gpurun_out/r2f_survival_steady.log
