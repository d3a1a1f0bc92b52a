#!/bin/bash
# tools/ncu_export.sh REP   -- on the GPU box: turn one .ncu-rep into the CSV pages the summaries are made from, then delete it
# (gpurun brings back at most 64 MiB; a --set full report with sources is ~25 MiB, its CSV pages ~3 MiB).
rep=$1; base=${rep%.ncu-rep}
ncu -i $rep --page raw --csv > ${base}_raw.csv 2>/dev/null
ncu -i $rep --page source --csv --print-source cuda,sass > ${base}_src.csv 2>/dev/null
ncu -i $rep --page source --csv --print-source sass > ${base}_sass.csv 2>/dev/null
[ -s ${base}_raw.csv ] && rm -f $rep
ls -la ${base}_*.csv
