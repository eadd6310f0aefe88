#!/bin/bash
# usage: gpu_prof_k.sh <kernel-regex> : one full ncu capture of the matching kernel (after a plain run)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --profile > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$1 -s 3 -c 1 -o gpurun_out/prof_$1 python bench.py --steps 2 --warmup 3 --profile > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
