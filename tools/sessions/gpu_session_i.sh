#!/bin/bash
# GPU session I (round 2): fused tiling step (first run), whole suite, e2e breakdown, config timings.
mkdir -p gpurun_out
python -m crispr_bean_b200.build --force > gpurun_out/r2i_build.log 2>&1 || { tail -30 gpurun_out/r2i_build.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_tiling_fused.py -q --tb=short 2>&1 | tail -60 > gpurun_out/r2i_tiling_tests.log
grep -E "passed|failed|FAILED|Error|errs|torch.float" gpurun_out/r2i_tiling_tests.log | head -30
timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -80 > gpurun_out/r2i_tests.log
grep -E "passed|failed|FAILED" gpurun_out/r2i_tests.log | tail
timeout 200 python tools/e2e_breakdown.py > gpurun_out/r2i_e2e_breakdown.json 2> gpurun_out/r2i_e2e_breakdown.err; cat gpurun_out/r2i_e2e_breakdown.json
timeout 300 python profiles/config_timings.py > gpurun_out/r2i_config_timings.jsonl 2> gpurun_out/r2i_config_timings.err; cat gpurun_out/r2i_config_timings.jsonl
timeout 200 python tools/quick_phase.py --tag r2i > gpurun_out/r2i_phase.json 2> gpurun_out/r2i_phase.err; cat gpurun_out/r2i_phase.json
