#!/bin/bash
# GPU session S (round 2, 4 GPUs): the device-side exchange against the host exchange at N = 4, weak scaling (no L2 flush, so
# the steps run 100 per call); steady-state step times per 100 steps of a sharded run.
mkdir -p gpurun_out
for mode in peer host; do
  if [ $mode = host ]; then export BEAN_NO_PEER_EXCHANGE=1; else unset BEAN_NO_PEER_EXCHANGE; fi
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 4 --steps 100 --warmup 5 --workload c4_survival --scaling weak --no-cpu-baseline \
      > gpurun_out/r2s_bench_c4_n4_weak_$mode.json 2> gpurun_out/r2s_bench_c4_n4_weak_$mode.err
  tail -c 300 gpurun_out/r2s_bench_c4_n4_weak_$mode.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2s_bench_c4_n4_weak_$mode.json').read().strip().splitlines()[-1])
print('c4 weak n4 $mode', {k:d[k] for k in ('ms_per_step','value')}, 'full run s', d['full_run']['seconds'], d.get('peer_exchange'), d['clocks'])
PY
done
