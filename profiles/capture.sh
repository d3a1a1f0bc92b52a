#!/bin/bash
# profiles/capture.sh TAG  -- run on the GPU box (through gpurun) AFTER a plain bench run has exited 0.
# Writes gpurun_out/launches_TAG.csv (per-launch durations of a short bench run) and gpurun_out/prof_TAG_guide.ncu-rep
# (one --set full capture of the dominant kernel, svi_guide_kernel, with source correlation).
TAG=${1:-r1}
set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:svi_guide_kernel -s 306 -c 1 -f -o gpurun_out/prof_${TAG}_guide \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2_$TAG.log 2>&1
ls -la gpurun_out/
