#!/bin/bash
# GPU session A (round 2): parity suite, fp32 error report, per-kernel times for three builds of the guide kernel.
mkdir -p gpurun_out
# the in-tree .so may be older than the sources (edits continue while the call waits for a GPU slot): build first
python -m crispr_bean_b200.build --force > gpurun_out/r2a_build.log 2>&1 || { tail -30 gpurun_out/r2a_build.log; exit 1; }
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2a_tests.log
timeout 200 python tools/fp32_error_report.py > gpurun_out/r2a_fp32_errors.json 2> gpurun_out/r2a_fp32_errors.err
timeout 200 python tools/quick_phase.py --tag packed8 > gpurun_out/r2a_phase_packed8.json 2> gpurun_out/r2a_phase_packed8.err
BEAN_NVCC_FLAGS=-DBEAN_GUIDE_MIN_CTAS=6 python -m crispr_bean_b200.build --force && timeout 200 python tools/quick_phase.py --tag packed6 > gpurun_out/r2a_phase_packed6.json 2> gpurun_out/r2a_phase_packed6.err
BEAN_NVCC_FLAGS=-DBEAN_NO_PACKED_FP32 python -m crispr_bean_b200.build --force && timeout 200 python tools/quick_phase.py --tag scalar8 > gpurun_out/r2a_phase_scalar8.json 2> gpurun_out/r2a_phase_scalar8.err
cat gpurun_out/r2a_phase_*.json
tail -8 gpurun_out/r2a_tests.log
