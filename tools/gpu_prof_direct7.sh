#!/bin/bash
# ncu --set full of the direct kernel at K = 7 (its default use): the A/B script launches it 10x at K = 8 first
mkdir -p gpurun_out
FRISK_AB_ONLY=direct python tools/direct_vs_bucket.py > gpurun_out/plain_direct7.log 2>&1 &&
FRISK_AB_ONLY=direct ncu --set full --clock-control none --import-source on -k regex:score_windows_direct -s 13 -c 1 -f -o gpurun_out/prof_direct7 python tools/direct_vs_bucket.py > gpurun_out/ncu_full_direct7.log 2>&1
tail -2 gpurun_out/ncu_full_direct7.log; cat gpurun_out/plain_direct7.log | head -3
