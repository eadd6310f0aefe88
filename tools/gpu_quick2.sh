#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 --profile > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','stage_ms','windows_per_s')})
except Exception as e:
    print("bench failed", e); print(open('gpurun_out/bench_quick.err').read()[-2000:])
PY
