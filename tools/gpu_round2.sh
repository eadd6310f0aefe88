#!/bin/bash
# Round-1 closing session: whole GPU suite, smoke, both bench arms, ncu launch list + full captures
# of the two hot kernels (each after a plain run of the same command).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 5000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 1500 gpurun_out/bench_ref.json
nproc
python bench.py --steps 2 --warmup 3 --profile > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --profile > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 2 --warmup 3 --profile > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_windows -s 3 -c 1 -o gpurun_out/prof_score python bench.py --steps 2 --warmup 3 --profile > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:bg_count -s 3 -c 1 -o gpurun_out/prof_bg_count python bench.py --steps 2 --warmup 3 --profile > gpurun_out/ncu_full_bg.log 2>&1
tail -2 gpurun_out/ncu_full_bg.log
ls -la gpurun_out
