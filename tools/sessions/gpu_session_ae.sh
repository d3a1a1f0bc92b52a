#!/bin/bash
# GPU session AE (round 2, 8 GPUs): the default bench line on 8 GPUs with the final kernels.
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 100 --warmup 5 \
    > gpurun_out/r2ae_bench_c5_n8.json 2> gpurun_out/r2ae_bench_c5_n8.err
tail -c 300 gpurun_out/r2ae_bench_c5_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ae_bench_c5_n8.json').read().strip().splitlines()[-1]); r=d['roofline']
print('c5 n8', {k:d[k] for k in ('ms_per_step','value','scaling')}, 'e2e', d['e2e']['value'], d['e2e'].get('ms_total_of_each_run'), 'full', d['full_run']['seconds'], 'kernels', r['ms_per_launch'], r['ms_per_launch_alpha_kernel'], r['ms_per_launch_variant_kernel'])
PY
