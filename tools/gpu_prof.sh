#!/bin/bash
# one full ncu capture of the score kernel + launch list (after a plain run of the same command)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --profile > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --profile > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 2 --warmup 3 --profile > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_windows -s 3 -c 1 -o gpurun_out/prof_score python bench.py --steps 2 --warmup 3 --profile > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
