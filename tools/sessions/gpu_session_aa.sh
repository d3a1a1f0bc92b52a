#!/bin/bash
# GPU session AA (round 2): per-variant kernel geometry (CTA size, lanes per variant) and programmatic dependent launch, A/B at the
# full and the quarter size; the sorting parity tests on the PDL build.
mkdir -p gpurun_out
for w in c5_genome_scale c5_quarter; do
  timeout 200 python tools/quick_phase.py --tag main --workload $w >> gpurun_out/r2aa_variants.jsonl 2>> gpurun_out/r2aa_variants.err
  for v in var512 lanes4 var512l4 pdl; do
    BEAN_B200_LIB=$PWD/crispr_bean_b200/variants/libbean_b200_$v.so timeout 200 python tools/quick_phase.py --tag $v --workload $w >> gpurun_out/r2aa_variants.jsonl 2>> gpurun_out/r2aa_variants.err
  done
done
cat gpurun_out/r2aa_variants.jsonl; tail -3 gpurun_out/r2aa_variants.err
BEAN_B200_LIB=$PWD/crispr_bean_b200/variants/libbean_b200_pdl.so timeout 600 python -m pytest tests/test_gpu_svi.py tests/test_gpu_golden.py tests/test_gpu_edge.py -q --tb=short 2>&1 | tail -8
