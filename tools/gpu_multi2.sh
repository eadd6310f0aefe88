#!/bin/bash
# 2-GPU session: the 2-rank test (both sharding schemes, NCCL and fused peer sum), bench at N=2 both ways
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_dist_gpu.py -x -q 2>&1 | tail -15
for mode in "" "--nccl"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 3 $mode > gpurun_out/bench_n2$mode.json 2> gpurun_out/bench_n2$mode.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n2$mode.json').read().strip().splitlines()[-1])
    print('$mode', {k:d[k] for k in ('value','ms_per_step','stage_ms')}, d['config']['parallelism'], d['e2e']['value'])
except Exception as e:
    print('bench failed', e)
PY
tail -4 gpurun_out/bench_n2$mode.err
done
