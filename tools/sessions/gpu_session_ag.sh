#!/bin/bash
# GPU session AG (round 2, 2 GPUs): the multi-GPU tests on the final library (one-thread-per-variant kernel, programmatic dependent
# launch, bean_row_const).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q --tb=short 2>&1 | tail -25 > gpurun_out/r2ag_multi_tests.log; tail -12 gpurun_out/r2ag_multi_tests.log
