#!/bin/bash
# GPU session AH (round 2): the direct test of bean_row_const.
timeout 300 python -m pytest tests/test_gpu_row_const.py -q --tb=short 2>&1 | tail -12
