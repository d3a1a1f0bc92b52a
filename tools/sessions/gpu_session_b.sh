#!/bin/bash
# GPU session B (round 2): failing tests with full output, whole suite, ncu capture of the guide / alpha kernels.
mkdir -p gpurun_out
python -m crispr_bean_b200.build --force > gpurun_out/r2b_build.log 2>&1 || { tail -30 gpurun_out/r2b_build.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_dirichlet.py "tests/test_gpu_configs.py::test_c3_tiling_shape" "tests/test_gpu_golden.py::test_fused_step_equals_reference_programs" -q -x --tb=short 2>&1 | tail -60 > gpurun_out/r2b_failing.log
timeout 300 python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -30 > gpurun_out/r2b_tests.log
timeout 600 bash profiles/capture.sh r2b > gpurun_out/r2b_capture.log 2>&1
tail -12 gpurun_out/r2b_tests.log
cat gpurun_out/steady_r2b.log | tail -3
