#!/bin/bash
# GPU session AB (round 2): more per-variant kernel geometries, with and without programmatic dependent launch.
mkdir -p gpurun_out
for w in c5_genome_scale c5_quarter; do
  for v in l2 l2t512 l1 l4t512pdl l2t512pdl; do
    BEAN_B200_LIB=$PWD/crispr_bean_b200/variants/libbean_b200_$v.so timeout 200 python tools/quick_phase.py --tag $v --workload $w >> gpurun_out/r2ab_variants.jsonl 2>> gpurun_out/r2ab_variants.err
  done
done
cat gpurun_out/r2ab_variants.jsonl; tail -3 gpurun_out/r2ab_variants.err
BEAN_B200_LIB=$PWD/crispr_bean_b200/variants/libbean_b200_l4t512pdl.so timeout 600 python -m pytest tests/test_gpu_svi.py tests/test_gpu_golden.py tests/test_gpu_survival_fused.py tests/test_gpu_tiling_fused.py tests/test_gpu_configs.py -q --tb=short 2>&1 | tail -6
