#!/bin/bash
# GPU session Y (round 2, 2 GPUs): the whole GPU suite including the multi-GPU tests.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r2y_tests.log; tail -12 gpurun_out/r2y_tests.log
