#!/bin/bash
# 2-GPU job: data-parallel equivalence test + N=2 bench (train-step scaling, allreduce_ms)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L
echo "== dp tests"; timeout 900 python -m pytest tests/test_gpu_dp.py -q -s -m gpu > $O/r2_dp2.log 2>&1; tail -15 $O/r2_dp2.log
echo "== bench N=2"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2_bench_n2.json 2> $O/r2_bench_n2.err; cat $O/r2_bench_n2.json | cut -c1-3000; tail -3 $O/r2_bench_n2.err | cut -c1-600
echo "== bench N=2 no overlap"; DASR_B200_DP_OVERLAP=0 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $O/r2_bench_n2_nooverlap.json 2> /dev/null; python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_n2.json', 'gpurun_out/r2_bench_n2_nooverlap.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'train', d['train']['value'], 'it/s', d['train']['ms_per_step'], 'ms allreduce_ms', d['train']['allreduce_ms'], '| fwd', d['value'])
    except Exception as e:
        print(f, 'unreadable', e)
PY
