#!/bin/bash
# full ncu capture of the direct score kernel (after a plain run of the same command)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --profile > gpurun_out/plain_direct.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_windows_direct -s 3 -c 1 -f -o gpurun_out/prof_direct python bench.py --steps 2 --warmup 3 --profile > gpurun_out/ncu_full_direct.log 2>&1
tail -2 gpurun_out/ncu_full_direct.log
