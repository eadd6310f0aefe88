#!/bin/bash
# ingest tests first (new code), then the whole GPU suite, smoke and a bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ingest_gpu.py -x -q 2>&1 | tail -25
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_ingest_gpu.py 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 4000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
