#!/bin/bash
# GPU session V (round 2, 8 GPUs): the default bench line on all GPUs of the box (strong scaling, 125 k guides per GPU).
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 8 --steps 100 --warmup 5 \
    > gpurun_out/r2v_bench_c5_n8.json 2> gpurun_out/r2v_bench_c5_n8.err
tail -c 400 gpurun_out/r2v_bench_c5_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench_c5_n8.json').read().strip().splitlines()[-1]); r=d['roofline']
print('c5 n8', {k:d[k] for k in ('ms_per_step','value','scaling')}, 'e2e', d['e2e']['value'], d['e2e'].get('ms_total_of_each_run'), 'full', d['full_run']['seconds'], 'kernels', r['ms_per_launch'], r['ms_per_launch_alpha_kernel'], r['ms_per_launch_variant_kernel'])
PY
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 8 --steps 100 --warmup 5 --workload c4_survival --no-cpu-baseline \
    > gpurun_out/r2v_bench_c4_n8.json 2> gpurun_out/r2v_bench_c4_n8.err
tail -c 400 gpurun_out/r2v_bench_c4_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench_c4_n8.json').read().strip().splitlines()[-1])
print('c4 n8', {k:d[k] for k in ('ms_per_step','value','scaling')}, 'full', d['full_run']['seconds'], d.get('peer_exchange'))
PY
