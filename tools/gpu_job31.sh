#!/bin/bash
# 4-GPU sanity run of the bench contract under torchrun
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 10 --warmup 3 --train-steps 10 --no-cpu-baseline > $O/r2_bench_n4.json 2> $O/r2_bench_n4.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_n4.json').read().strip().splitlines()[-1])
    print('N=4 fwd', d['value'], 'MP/s', d['ms_per_step'], 'ms | train', d['train']['value'], 'it/s', d['train']['ms_per_step'], 'ms allreduce_ms', d['train']['allreduce_ms'], '| dsn', d['dsn']['value'])
except Exception as e:
    print('unreadable', e); print(open('gpurun_out/r2_bench_n4.err').read()[-2000:])
PY
