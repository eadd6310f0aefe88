#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for mode in "" "--nccl"; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 30 --warmup 5 $mode > gpurun_out/bench_n8$mode.json 2> gpurun_out/bench_n8$mode.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n8$mode.json').read().strip().splitlines()[-1])
    print('$mode', {k:d[k] for k in ('value','ms_per_step','stage_ms')}, d['config']['parallelism'], d['e2e']['value'])
except Exception as e:
    print('bench failed', e)
PY
tail -2 gpurun_out/bench_n8$mode.err
done
