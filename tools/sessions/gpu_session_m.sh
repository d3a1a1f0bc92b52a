#!/bin/bash
# GPU session M (round 2): suite after the replicate-parallel tiling kernel, tiling step time, config timings, final guide-kernel
# ncu capture (exported to CSV on the box).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r2m_tests.log
grep -E "passed|failed|FAILED" gpurun_out/r2m_tests.log | tail
timeout 120 python profiles/tiling_steady.py 300 > gpurun_out/r2m_tiling_steady.log 2>&1; cat gpurun_out/r2m_tiling_steady.log
timeout 200 python tools/quick_phase.py --tag main > gpurun_out/r2m_phase.json 2> gpurun_out/r2m_phase.err; cat gpurun_out/r2m_phase.json
timeout 300 python profiles/config_timings.py > gpurun_out/r2m_config_timings.jsonl 2> gpurun_out/r2m_config_timings.err; cat gpurun_out/r2m_config_timings.jsonl
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:tiling|svi_variant" -s 300 -c 30 --csv --log-file gpurun_out/launches_r2m_tiling.csv \
    python profiles/tiling_steady.py 200 > gpurun_out/r2m_ncu_tiling1.log 2>&1
for k in guide; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:svi_${k}_kernel -s 500 -c 1 -f -o gpurun_out/prof_r2m_$k \
      python profiles/steady_state.py 600 > gpurun_out/r2m_ncu_$k.log 2>&1
  bash tools/ncu_export.sh gpurun_out/prof_r2m_$k.ncu-rep
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tiling_guide_kernel -s 100 -c 1 -f -o gpurun_out/prof_r2m_tiling_guide \
    python profiles/tiling_steady.py 200 > gpurun_out/r2m_ncu_tiling2.log 2>&1
bash tools/ncu_export.sh gpurun_out/prof_r2m_tiling_guide.ncu-rep
du -sh gpurun_out
