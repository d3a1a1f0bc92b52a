#!/bin/bash
# GPU session X (round 2, 2 GPUs): the tiling step sharded over two GPUs (guide blocks, per-edit gradient sums all-reduced each
# step), alone and through run_inference; the other multi-GPU tests; the single-GPU tiling tests after the kernel change.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_tiling_fused.py tests/test_gpu_survival_fused.py -q --tb=short 2>&1 | tail -40 > gpurun_out/r2x_tests.log; tail -30 gpurun_out/r2x_tests.log
