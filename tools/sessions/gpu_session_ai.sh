#!/bin/bash
# GPU session AI (round 2): the whole single-GPU suite on the final library, last check of the round.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -12 > gpurun_out/r2ai_tests.log; grep -E "passed|failed|FAILED" gpurun_out/r2ai_tests.log | tail -4
