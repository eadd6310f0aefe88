#!/bin/bash
# direct kernel bring-up: its parity tests, then the bench with both CTA sizes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "direct or low_complexity or window_tables or fresh_genomes or bit_reproducible" 2>&1 | tail -15
for nt in 256 384 bucket; do
opt="direct_threads=$nt"; [ $nt = bucket ] && opt="force_bucket_kernel=1"
timeout 300 python bench.py --steps 20 --warmup 3 --profile --option $opt > gpurun_out/bench_direct_$nt.json 2> gpurun_out/bench_direct_$nt.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_direct_$nt.json').read().strip().splitlines()[-1])
    print('$nt', {k:d[k] for k in ('value','ms_per_step','stage_ms','windows_per_s')}, d['config'].get('score_kernel_occupancy'))
except Exception as e:
    print("bench failed", e); print(open('gpurun_out/bench_direct_$nt.err').read()[-2000:])
PY
done
