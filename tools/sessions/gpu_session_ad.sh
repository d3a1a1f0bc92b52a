#!/bin/bash
# GPU session AD (round 2): the edge tests (long / short guide segments of the per-variant kernel) and the whole suite once more.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_edge.py -q --tb=short 2>&1 | tail -8
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r2ad_tests.log; grep -E "passed|failed|FAILED" gpurun_out/r2ad_tests.log | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-200
