#!/bin/bash
# 2-GPU session: the 2-rank test (both sharding schemes), bench at N=2
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_dist_gpu.py -x -q 2>&1 | tail -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 2500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
