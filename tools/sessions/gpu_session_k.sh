#!/bin/bash
# GPU session K (round 2): suite after the variant-kernel / alpha-kernel / guide-kernel changes, A/B of the Philox-key and
# occupancy builds, ncu captures exported to CSV on the box (gpurun_out must stay under 64 MiB), bench lines.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60 > gpurun_out/r2k_tests.log
grep -E "passed|failed|FAILED" gpurun_out/r2k_tests.log | tail
timeout 200 python tools/quick_phase.py --tag main >> gpurun_out/r2k_variants.jsonl 2>> gpurun_out/r2k_variants.err
for v in keys keys8 keys6; do
  BEAN_B200_LIB=$PWD/crispr_bean_b200/variants/libbean_b200_$v.so timeout 200 python tools/quick_phase.py --tag $v >> gpurun_out/r2k_variants.jsonl 2>> gpurun_out/r2k_variants.err
done
cat gpurun_out/r2k_variants.jsonl
timeout 300 python profiles/survival_steady.py 300 > gpurun_out/r2k_survival_steady.log 2>&1; cat gpurun_out/r2k_survival_steady.log
timeout 120 python profiles/tiling_steady.py 300 > gpurun_out/r2k_tiling_steady.log 2>&1; cat gpurun_out/r2k_tiling_steady.log
# ncu: launch list of a short bench run, then one --set full capture per kernel, each exported to CSV and deleted
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2k.csv \
    python bench.py --steps 3 --warmup 3 --burn-in 0 --no-cpu-baseline --full-run-steps 0 > gpurun_out/r2k_ncu_launches.log 2>&1
for k in guide alpha; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:svi_${k}_kernel -s 500 -c 1 -f -o gpurun_out/prof_r2k_$k \
      python profiles/steady_state.py 600 > gpurun_out/r2k_ncu_$k.log 2>&1
  bash tools/ncu_export.sh gpurun_out/prof_r2k_$k.ncu-rep
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:surv_guide_kernel -s 200 -c 1 -f -o gpurun_out/prof_r2k_surv_guide \
    python profiles/survival_steady.py 300 > gpurun_out/r2k_ncu_surv.log 2>&1
bash tools/ncu_export.sh gpurun_out/prof_r2k_surv_guide.ncu-rep
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:tiling|svi_variant" -s 300 -c 30 --csv --log-file gpurun_out/launches_r2k_tiling.csv \
    python profiles/tiling_steady.py 200 > gpurun_out/r2k_ncu_tiling1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tiling_guide_kernel -s 100 -c 1 -f -o gpurun_out/prof_r2k_tiling_guide \
    python profiles/tiling_steady.py 200 > gpurun_out/r2k_ncu_tiling2.log 2>&1
bash tools/ncu_export.sh gpurun_out/prof_r2k_tiling_guide.ncu-rep
timeout 600 python bench.py > gpurun_out/r2k_bench_c5.json 2> gpurun_out/r2k_bench_c5.err; tail -c 800 gpurun_out/r2k_bench_c5.json
timeout 300 python bench.py --workload c4_survival > gpurun_out/r2k_bench_c4.json 2> gpurun_out/r2k_bench_c4.err; tail -c 400 gpurun_out/r2k_bench_c4.json
du -sh gpurun_out; ls -la gpurun_out | tail -30
