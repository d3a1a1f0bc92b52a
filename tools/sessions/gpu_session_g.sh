#!/bin/bash
# GPU session G (round 2, N GPUs): multi-GPU tests + the strong-scaling bench lines.  usage: gpu_session_g.sh N
N=${1:-2}
mkdir -p gpurun_out
python -m crispr_bean_b200.build --force > gpurun_out/r2g_build.log 2>&1 || { tail -30 gpurun_out/r2g_build.log; exit 1; }
if [ "$N" = "2" ]; then
  timeout 300 python -m pytest tests/test_gpu_multi.py -q --tb=short 2>&1 | tail -40 > gpurun_out/r2g_multi_tests.log
  grep -E "passed|failed|FAILED|Error" gpurun_out/r2g_multi_tests.log | tail
fi
for wl in c5_genome_scale c4_survival; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --workload $wl > gpurun_out/r2g_bench_${wl}_n$N.json 2> gpurun_out/r2g_bench_${wl}_n$N.err
  echo "== $wl N=$N rc=$?"; head -c 900 gpurun_out/r2g_bench_${wl}_n$N.json; echo; tail -3 gpurun_out/r2g_bench_${wl}_n$N.err
done
