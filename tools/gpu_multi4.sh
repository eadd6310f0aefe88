#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; tail -c 1800 gpurun_out/bench_n4.json; tail -3 gpurun_out/bench_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29556 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 2>&1 | tail -c 600
