#!/bin/bash
# whole GPU suite, smoke, bench (own arm + reference arm)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 4500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python tools/ingest_prof.py 4 2>&1 | tail -3
