#!/bin/bash
# GPU session Q (round 2, 2 GPUs): the device-side exchange of the survival step's library-wide sums through CUDA-IPC peer memory.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -q --tb=short -k "fused or run_inference" 2>&1 | tail -40 > gpurun_out/r2q_multi_tests.log; tail -25 gpurun_out/r2q_multi_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 100 --warmup 5 --workload c4_survival --no-cpu-baseline \
    > gpurun_out/r2q_bench_c4_n2.json 2> gpurun_out/r2q_bench_c4_n2.err
tail -c 600 gpurun_out/r2q_bench_c4_n2.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q_bench_c4_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('ms_per_step','value')}, d['e2e']['value'], d['full_run']['seconds'])
PY
BEAN_NO_PEER_EXCHANGE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 100 --warmup 5 --workload c4_survival --no-cpu-baseline \
    > gpurun_out/r2q_bench_c4_n2_host.json 2> gpurun_out/r2q_bench_c4_n2_host.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q_bench_c4_n2_host.json').read().strip().splitlines()[-1])
print('host exchange', {k:d[k] for k in ('ms_per_step','value')}, d['e2e']['value'], d['full_run']['seconds'])
PY
