#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --profile > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_windows -s 3 -c 1 -o gpurun_out/prof_score python bench.py --steps 2 --warmup 3 --profile > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
