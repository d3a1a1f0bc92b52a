#!/bin/bash
# GPU session J (round 2): fixes of session I re-tested, whole suite, A/B of the layer-constant guide kernel builds, ncu captures
# of the staged guide kernel / alpha / survival / tiling kernels, bench lines.  The library travels prebuilt (no rebuild here).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tiling_fused.py "tests/test_gpu_e2e_reference_screens.py" -q --tb=short 2>&1 | tail -40 > gpurun_out/r2j_fix_tests.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2j_fix_tests.log | head
timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60 > gpurun_out/r2j_tests.log
grep -E "passed|failed|FAILED" gpurun_out/r2j_tests.log | tail
for v in base hoist unroll both; do
  BEAN_B200_LIB=$PWD/crispr_bean_b200/variants/libbean_b200_$v.so timeout 200 python tools/quick_phase.py --tag $v >> gpurun_out/r2j_variants.jsonl 2>> gpurun_out/r2j_variants.err
done
cat gpurun_out/r2j_variants.jsonl
timeout 700 bash profiles/capture.sh r2j > gpurun_out/r2j_capture.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:surv_guide_kernel -s 200 -c 1 -f -o gpurun_out/prof_r2j_surv_guide python profiles/survival_steady.py 300 > gpurun_out/r2j_ncu_surv.log 2>&1
timeout 120 python profiles/tiling_steady.py 300 > gpurun_out/r2j_tiling_steady.log 2>&1; cat gpurun_out/r2j_tiling_steady.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 30 --csv --log-file gpurun_out/launches_r2j_tiling.csv python profiles/tiling_steady.py 200 > gpurun_out/r2j_ncu_tiling1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tiling_guide_kernel -s 100 -c 1 -f -o gpurun_out/prof_r2j_tiling_guide python profiles/tiling_steady.py 200 > gpurun_out/r2j_ncu_tiling2.log 2>&1
timeout 600 python bench.py > gpurun_out/r2j_bench_c5.json 2> gpurun_out/r2j_bench_c5.err; tail -c 1500 gpurun_out/r2j_bench_c5.json
timeout 300 python bench.py --workload c4_survival > gpurun_out/r2j_bench_c4.json 2> gpurun_out/r2j_bench_c4.err; tail -c 600 gpurun_out/r2j_bench_c4.json
ls -la gpurun_out | tail -15
