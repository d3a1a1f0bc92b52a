#!/bin/bash
# GPU session P (round 2): where the var_mini fp32 alpha_pi error comes from; alpha-kernel CTA size A/B at full and quarter size.
mkdir -p gpurun_out
timeout 300 python tools/diag_var_mini.py > gpurun_out/r2p_diag_var_mini.jsonl 2> gpurun_out/r2p_diag_var_mini.err; tail -c 300 gpurun_out/r2p_diag_var_mini.err; cut -c1-1500 gpurun_out/r2p_diag_var_mini.jsonl
for w in c5_genome_scale c5_quarter; do
  timeout 200 python tools/quick_phase.py --tag main --workload $w >> gpurun_out/r2p_variants.jsonl 2>> gpurun_out/r2p_variants.err
  for v in alpha64 alpha32; do
    BEAN_B200_LIB=$PWD/crispr_bean_b200/variants/libbean_b200_$v.so timeout 200 python tools/quick_phase.py --tag $v --workload $w >> gpurun_out/r2p_variants.jsonl 2>> gpurun_out/r2p_variants.err
  done
done
cat gpurun_out/r2p_variants.jsonl; tail -3 gpurun_out/r2p_variants.err
