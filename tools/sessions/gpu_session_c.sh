#!/bin/bash
# GPU session C (round 2): whole suite (fused survival step, specialised guide kernels), per-kernel times, fp32 error report.
mkdir -p gpurun_out
python -m crispr_bean_b200.build --force > gpurun_out/r2c_build.log 2>&1 || { tail -30 gpurun_out/r2c_build.log; exit 1; }
timeout 400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -150 > gpurun_out/r2c_tests.log
timeout 200 python tools/fp32_error_report.py > gpurun_out/r2c_fp32_errors.json 2> gpurun_out/r2c_fp32_errors.err
timeout 200 python tools/quick_phase.py --tag fast8 > gpurun_out/r2c_phase_fast8.json 2> gpurun_out/r2c_phase_fast8.err
BEAN_NVCC_FLAGS=-DBEAN_GUIDE_MIN_CTAS=6 python -m crispr_bean_b200.build --force && timeout 200 python tools/quick_phase.py --tag fast6 > gpurun_out/r2c_phase_fast6.json 2> gpurun_out/r2c_phase_fast6.err
cat gpurun_out/r2c_phase_*.json
grep -E "passed|failed|FAILED" gpurun_out/r2c_tests.log | tail -25
