#!/bin/bash
# GPU session R (round 2, 4 GPUs): suite incl. the multi-GPU tests, smoke(), weak-scaling lines at N = 4, the device-side
# exchange against the host exchange where the step is short (c4 at its real size, 2 GPUs).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q --tb=short 2>&1 | tail -30 > gpurun_out/r2r_tests.log; grep -E "passed|failed|FAILED" gpurun_out/r2r_tests.log | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.log 2>&1; tail -4 gpurun_out/r2r_smoke.log
for w in c5_genome_scale c4_survival; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 4 --steps 100 --warmup 5 --workload $w --scaling weak --no-cpu-baseline \
      > gpurun_out/r2r_bench_${w}_n4_weak.json 2> gpurun_out/r2r_bench_${w}_n4_weak.err
  tail -c 300 gpurun_out/r2r_bench_${w}_n4_weak.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2r_bench_${w}_n4_weak.json').read().strip().splitlines()[-1])
print('$w weak n4', {k:d[k] for k in ('ms_per_step','value','scaling')}, d['e2e']['value'], d['full_run']['seconds'])
PY
done
for mode in peer host; do
  if [ $mode = host ]; then export BEAN_NO_PEER_EXCHANGE=1; else unset BEAN_NO_PEER_EXCHANGE; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29583 bench.py --gpus 2 --steps 100 --warmup 5 --workload c4_survival_small --no-cpu-baseline \
      > gpurun_out/r2r_bench_c4small_n2_$mode.json 2> gpurun_out/r2r_bench_c4small_n2_$mode.err
  tail -c 300 gpurun_out/r2r_bench_c4small_n2_$mode.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2r_bench_c4small_n2_$mode.json').read().strip().splitlines()[-1])
print('c4 small n2 $mode', {k:d[k] for k in ('ms_per_step','value')}, 'full run s', d['full_run']['seconds'])
PY
done
unset BEAN_NO_PEER_EXCHANGE
timeout 300 python bench.py --workload c4_survival_small --no-cpu-baseline > gpurun_out/r2r_bench_c4small_n1.json 2> gpurun_out/r2r_bench_c4small_n1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2r_bench_c4small_n1.json').read().strip().splitlines()[-1])
print('c4 small n1', {k:d[k] for k in ('ms_per_step','value')}, 'full run s', d['full_run']['seconds'])
PY
