#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scale_gpu.py -x -q 2>&1 | tail -30
python tools/ingest_prof.py 4 2>&1 | tail -5
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/ingest_launches.csv python tools/ingest_prof.py 2 > gpurun_out/ingest_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/ingest_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-40:]:
    print(r[4][:60], r[-1])
PY
