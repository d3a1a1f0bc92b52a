#!/bin/bash
# GPU session D (round 2): suite, fp32 errors, survival scale, config timings, the default bench + reference arm.
mkdir -p gpurun_out
python -m crispr_bean_b200.build --force > gpurun_out/r2d_build.log 2>&1 || { tail -30 gpurun_out/r2d_build.log; exit 1; }
timeout 400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -100 > gpurun_out/r2d_tests.log
timeout 200 python tools/fp32_error_report.py > gpurun_out/r2d_fp32_errors.json 2> gpurun_out/r2d_fp32_errors.err
timeout 300 python profiles/survival_scale.py 700 20000 200000 > gpurun_out/r2d_survival_scale.jsonl 2> gpurun_out/r2d_survival_scale.err
timeout 300 python profiles/config_timings.py > gpurun_out/r2d_config_timings.jsonl 2> gpurun_out/r2d_config_timings.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2d_bench_ref.json 2> gpurun_out/r2d_bench_ref.err
timeout 600 python bench.py --workload c4_survival --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench_c4.json 2> gpurun_out/r2d_bench_c4.err
grep -E "passed|failed|FAILED" gpurun_out/r2d_tests.log | tail
cat gpurun_out/r2d_survival_scale.jsonl gpurun_out/r2d_config_timings.jsonl
head -c 1500 gpurun_out/r2d_bench.json; echo; head -c 600 gpurun_out/r2d_bench_ref.json; echo; head -c 800 gpurun_out/r2d_bench_c4.json
tail -3 gpurun_out/r2d_bench.err gpurun_out/r2d_bench_c4.err gpurun_out/r2d_bench_ref.err
