#!/bin/bash
# GPU session H (round 2): suite with the cp.async-staged guide kernel; A/B staged vs direct loads, 8 vs 7 CTAs per SM.
mkdir -p gpurun_out
python -m crispr_bean_b200.build --force > gpurun_out/r2h_build.log 2>&1 || { tail -30 gpurun_out/r2h_build.log; exit 1; }
timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60 > gpurun_out/r2h_tests.log
grep -E "passed|failed|FAILED" gpurun_out/r2h_tests.log | tail
timeout 200 python tools/quick_phase.py --tag staged8 > gpurun_out/r2h_phase_staged8.json 2> gpurun_out/r2h_phase_staged8.err
BEAN_NVCC_FLAGS=-DBEAN_GUIDE_MIN_CTAS=7 python -m crispr_bean_b200.build --force && timeout 200 python tools/quick_phase.py --tag staged7 > gpurun_out/r2h_phase_staged7.json 2> gpurun_out/r2h_phase_staged7.err
BEAN_NVCC_FLAGS=-DBEAN_GUIDE_MIN_CTAS=6 python -m crispr_bean_b200.build --force && timeout 200 python tools/quick_phase.py --tag staged6 > gpurun_out/r2h_phase_staged6.json 2> gpurun_out/r2h_phase_staged6.err
BEAN_NVCC_FLAGS=-DBEAN_NO_STAGE python -m crispr_bean_b200.build --force && timeout 200 python tools/quick_phase.py --tag direct8 > gpurun_out/r2h_phase_direct8.json 2> gpurun_out/r2h_phase_direct8.err
cat gpurun_out/r2h_phase_*.json
