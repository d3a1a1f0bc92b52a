#!/bin/bash
# profiles/capture.sh TAG  -- run on the GPU box (through gpurun).  Each ncu pass only after the same command exited 0 plain.
#   gpurun_out/plain_TAG.json          the bench line of the plain run (no profiler)
#   gpurun_out/launches_TAG.csv        per-launch durations of a short bench run (ncu --metrics gpu__time_duration.sum)
#   gpurun_out/prof_TAG_guide.ncu-rep  one --set full capture of svi_guide_kernel in the STEADY STATE (step 500 of a run)
#   gpurun_out/prof_TAG_alpha.ncu-rep  the same for svi_alpha_kernel
TAG=${1:-r1}
set -x
python bench.py --steps 3 --warmup 3 --burn-in 0 --no-cpu-baseline --full-run-steps 0 > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --burn-in 0 --no-cpu-baseline --full-run-steps 0 > gpurun_out/ncu1_$TAG.log 2>&1
python profiles/steady_state.py 600 > gpurun_out/steady_$TAG.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:svi_guide_kernel -s 500 -c 1 -f -o gpurun_out/prof_${TAG}_guide \
    python profiles/steady_state.py 600 > gpurun_out/ncu2_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:svi_alpha_kernel -s 500 -c 1 -f -o gpurun_out/prof_${TAG}_alpha \
    python profiles/steady_state.py 600 > gpurun_out/ncu3_$TAG.log 2>&1
ls -la gpurun_out/ | tail -8
