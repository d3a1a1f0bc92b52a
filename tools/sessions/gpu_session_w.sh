#!/bin/bash
# GPU session W (round 2): survival guide kernel occupancy A/B (5 / 6 / 7 / 8 CTAs per SM) with the inlined Philox.
mkdir -p gpurun_out
for v in surv6 surv5 surv7 surv8; do
  echo "== $v" >> gpurun_out/r2w_survival_variants.log
  BEAN_B200_LIB=$PWD/crispr_bean_b200/variants/libbean_b200_$v.so timeout 200 python profiles/survival_steady.py 300 >> gpurun_out/r2w_survival_variants.log 2>&1
done
cat gpurun_out/r2w_survival_variants.log
