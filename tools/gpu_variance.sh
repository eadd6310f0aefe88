#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=uuid --format=csv,noheader | tail -c 14
for i in 1 2 3; do
python bench.py --steps 20 --warmup 3 --profile | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('score_ms', round(d['stage_ms']['score'],4), d['config']['score_kernel_occupancy'], 'power', d['clocks']['power_w_max'])"
done
