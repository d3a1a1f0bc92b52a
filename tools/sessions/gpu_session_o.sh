#!/bin/bash
# GPU session O (round 2, 4 GPUs): multi-GPU tests, strong-scaling bench lines at N = 2 and 4 (c5) and N = 4 (c4 survival: the
# abundance all-reduce inside the loop).
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_multi.py -q --tb=short 2>&1 | tail -15 > gpurun_out/r2o_multi_tests.log; cat gpurun_out/r2o_multi_tests.log | tail -5
for n in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 100 --warmup 5 \
      > gpurun_out/r2o_bench_c5_n$n.json 2> gpurun_out/r2o_bench_c5_n$n.err
  tail -c 400 gpurun_out/r2o_bench_c5_n$n.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2o_bench_c5_n$n.json').read().strip().splitlines()[-1])
print($n, {k:d[k] for k in ('ms_per_step','value','scaling')}, d['e2e']['value'], d.get('full_run',{}).get('seconds'))
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 100 --warmup 5 --workload c4_survival \
    > gpurun_out/r2o_bench_c4_n4.json 2> gpurun_out/r2o_bench_c4_n4.err
tail -c 400 gpurun_out/r2o_bench_c4_n4.err; tail -c 700 gpurun_out/r2o_bench_c4_n4.json
du -sh gpurun_out
